#!/bin/bash
# One GPU-box visit: parity tests, headline bench, kernel launch list, one full ncu capture, tuning sweeps.
# Everything lands in gpurun_out/ (scratch); summaries worth keeping are copied to profiles/ afterwards.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== nproc $(nproc), mem $(free -g | awk '/Mem/{print $2}') GiB"; nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv,noheader
echo "== pytest gpu"
timeout 900 python -m pytest tests -x -q -m gpu -s 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "== bench (default: 4 GiB)"
timeout 600 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench.json; tail -5 gpurun_out/bench_err.log
echo "== bench reference arm"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/bench_err.log | tee gpurun_out/bench_ref.json
echo "== sweeps (1 GiB, 3 steps)"
for cfg in "13 11 0" "13 12 0" "14 12 0" "14 13 0" "13 11 8" "13 11 12"; do
  set -- $cfg
  echo "-- piece_log2=$1 hash_bits=$2 warps=$3"
  QZB200_PIECE_LOG2=$1 QZB200_HASH_BITS=$2 QZB200_WARPS=$3 QZ_BENCH_NOCPU=1 timeout 300 python bench.py --steps 3 --warmup 2 --gib 1 2>>gpurun_out/bench_err.log | tee -a gpurun_out/sweep.jsonl
done
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 1 --gib 0.5 > gpurun_out/ncu_launch_run.log 2>&1; tail -3 gpurun_out/ncu_launch_run.log
echo "== ncu full on the deflate kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qzb_deflate_pieces -s 2 -c 1 -o gpurun_out/prof_deflate -f \
   python bench.py --steps 1 --warmup 1 --gib 0.5 > gpurun_out/ncu_full_run.log 2>&1; tail -3 gpurun_out/ncu_full_run.log
ls -la gpurun_out | head -30
