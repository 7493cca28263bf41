#!/bin/bash
# round 2, first contact of the window kernel: parity suite, geometry sweep, quick bench
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
echo "== nproc $(nproc)"; nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
echo "== pytest gpu"; timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/r02a_pytest_gpu.log
echo "== geometry sweep (window kernel: groups x units x table entries)"
for cfg in "4 2 2048" "3 2 2048" "2 2 2048" "4 2 1024" "4 2 1536" "3 2 2560" "4 1 2048" "2 1 2048"; do set -- $cfg
  echo -n "groups $1 units $2 tent $3: "; QZB200_WINDOW_GROUPS=$1 QZB200_WINDOW_UNITS=$2 QZB200_WINDOW_TENT=$3 timeout 120 python tools/gpu_geom.py 2>&1 | tail -1; done | tee gpurun_out/r02a_window_geometry.log
echo -n "per-piece kernel: "; QZB200_WINDOW=0 timeout 120 python tools/gpu_geom.py 2>&1 | tail -1 | tee -a gpurun_out/r02a_window_geometry.log
echo "== bench"; QZ_BENCH_NOCPU=1 timeout 600 python bench.py --steps 3 > gpurun_out/r02a_bench_quick.json 2> gpurun_out/bench_err.log; cat gpurun_out/r02a_bench_quick.json; tail -3 gpurun_out/bench_err.log
