#!/bin/bash
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
echo "== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench.json; tail -3 gpurun_out/bench_err.log
echo "== geometry sweep (2 GiB)"
for cfg in "24 12" "22 13" "24 10" "20 12" "26 11" "16 16"; do set -- $cfg; echo "-- warps=$1 buffers=$2"
  QZB200_WARPS=$1 QZB200_BUFFERS=$2 QZ_BENCH_NOCPU=1 timeout 300 python bench.py --steps 3 --warmup 2 --gib 2 2>>gpurun_out/bench_err.log | tee -a gpurun_out/sweep.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['roofline']['ms_per_launch'])"; done
echo "== extra perf"; timeout 900 python tools/gpu_perf_extra.py 2>gpurun_out/extra_err.log | tee gpurun_out/extra.json; tail -3 gpurun_out/extra_err.log
