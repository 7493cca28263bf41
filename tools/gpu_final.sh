#!/bin/bash
# Round-end evidence on one GPU box: parity tests, headline bench + reference arm, ncu launch list, full ncu
# captures (deflate piece kernel incl. DRAM bytes, inflate kernel), secondary measurements.
# Everything lands in gpurun_out/ (scratch); tools/collect_profiles.py copies the summaries into profiles/.
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
echo "== nproc $(nproc)"; nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv,noheader
echo "== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench_err.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench_err.log
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/bench_err.log | tee gpurun_out/bench_ref.json
echo "== hash-table sizes of the group kernel (buffers fitted to the 227 KB)"
for hb in 10 9; do echo -n "hash bits $hb: "; QZB200_GROUP_HASH_BITS=$hb timeout 120 python tools/gpu_geom.py 2>&1 | tail -1; done | tee gpurun_out/group_hash_bits.log
echo "== bench variants (e2e only matters): tail taper, two submitting threads"
QZ_BENCH_NOCPU=1 QZB200_TAPER=1 timeout 300 python bench.py --steps 3 > gpurun_out/bench_taper.json 2>> gpurun_out/bench_err.log; python -c "import json; b=json.load(open('gpurun_out/bench_taper.json')); print('taper', b['value'], b['e2e']['value'])"
QZ_BENCH_NOCPU=1 timeout 300 python bench.py --steps 3 --e2e-threads 2 > gpurun_out/bench_2thr.json 2>> gpurun_out/bench_err.log; python -c "import json; b=json.load(open('gpurun_out/bench_2thr.json')); print('2 threads', b['value'], b['e2e']['value'], 'one thread', b['e2e']['one_thread']['value'])"
echo "== ncu launch list"
QZ_BENCH_NOCPU=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 1 --gib 0.5 > gpurun_out/ncu_launch_run.log 2>&1; tail -1 gpurun_out/ncu_launch_run.log
echo "== ncu full: deflate kernel"
QZ_BENCH_NOCPU=1 timeout 900 ncu --set full --clock-control none --import-source on -k "regex:qzb_deflate_(groups|pieces)" -s 1 -c 1 -o gpurun_out/prof_deflate -f \
   python bench.py --steps 1 --warmup 1 --gib 0.5 > gpurun_out/ncu_full_run.log 2>&1; tail -1 gpurun_out/ncu_full_run.log
echo "== ncu full: inflate kernel"
INFL_MIB=512 INFL_REPS=2 INFL_CASES=ours timeout 600 ncu --set full --clock-control none --import-source on -k regex:qzb_inflate_kernel -s 1 -c 1 -o gpurun_out/prof_inflate -f python tools/gpu_inflate_bench.py > gpurun_out/ncu_inflate.log 2>&1; tail -1 gpurun_out/ncu_inflate.log
echo "== inflate bench"; INFL_MIB=2048 INFL_REF_MIB=2048 timeout 900 python tools/gpu_inflate_bench.py 2>&1 | tail -1 | tee gpurun_out/inflate_bench.json
echo "== extra"; EXTRA_STREAM_MIB=256 timeout 900 python tools/gpu_perf_extra.py > gpurun_out/extra.json 2> gpurun_out/extra_err.log; cat gpurun_out/extra.json; tail -3 gpurun_out/extra_err.log
