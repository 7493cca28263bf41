#!/bin/bash
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
echo "== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench.json; tail -3 gpurun_out/bench_err.log
echo "== batch sweep"
for mb in 64 256; do echo "-- batch=$mb"; QZB200_BATCH_MB=$mb QZ_BENCH_NOCPU=1 timeout 300 python bench.py --steps 3 --warmup 2 --gib 2 2>>gpurun_out/bench_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'])"; done
