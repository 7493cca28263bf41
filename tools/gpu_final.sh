#!/bin/bash
# Round-end evidence on one GPU box: parity tests, headline bench with its secondary legs + reference arm, ncu launch list,
# full ncu captures (deflate window kernel incl. DRAM bytes, LZ4 window kernel, inflate kernel), per-phase shares.
# Everything lands in gpurun_out/ (scratch); tools/collect_profiles.py copies the summaries into profiles/.
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
echo "== nproc $(nproc)"; nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv,noheader
echo "== pytest gpu"; timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
echo "== bench"; timeout 1500 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench_err.log; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); s=d.pop('secondary'); print(json.dumps(d)); print(json.dumps(s, indent=1))"; tail -3 gpurun_out/bench_err.log
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/bench_err.log | tee gpurun_out/bench_ref.json
echo "== phases"; timeout 200 python tools/gpu_phases.py 2>&1 | tail -1 | tee gpurun_out/phases_window.json
echo "== ncu launch list"
QZ_BENCH_NOCPU=1 QZ_BENCH_SECONDARY=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 1 --gib 0.5 > gpurun_out/ncu_launch_run.log 2>&1; tail -1 gpurun_out/ncu_launch_run.log
echo "== ncu full: deflate window kernel"
QZ_BENCH_NOCPU=1 QZ_BENCH_SECONDARY=0 timeout 900 ncu --set full --clock-control none --import-source on -k "regex:qzb_deflate_(window|pieces)" -s 1 -c 1 -o gpurun_out/prof_deflate -f \
   python bench.py --steps 1 --warmup 1 --gib 0.5 > gpurun_out/ncu_full_run.log 2>&1; tail -1 gpurun_out/ncu_full_run.log
echo "== ncu full: inflate kernel"
INFL_MIB=512 INFL_REPS=2 INFL_CASES=ours timeout 600 ncu --set full --clock-control none --import-source on -k regex:qzb_inflate_kernel -s 1 -c 1 -o gpurun_out/prof_inflate -f python tools/gpu_inflate_bench.py > gpurun_out/ncu_inflate.log 2>&1; tail -1 gpurun_out/ncu_inflate.log
echo "== ncu full: LZ4 window kernel"
EXTRA_MIB=512 EXTRA_NOCPU=1 EXTRA_STREAM_MIB=0 EXTRA_ONLY=lz4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:qzb_lz4_window_kernel -s 1 -c 1 -o gpurun_out/prof_lz4 -f python tools/gpu_perf_extra.py > gpurun_out/ncu_lz4.log 2>&1; tail -1 gpurun_out/ncu_lz4.log
echo "== extra"; EXTRA_STREAM_MIB=0 timeout 900 python tools/gpu_perf_extra.py > gpurun_out/extra.json 2> gpurun_out/extra_err.log; cat gpurun_out/extra.json; tail -3 gpurun_out/extra_err.log
echo "== pcie duplex"; timeout 300 python tools/gpu_pcie_duplex.py 2>&1 | tail -1 | tee gpurun_out/pcie_duplex.json
