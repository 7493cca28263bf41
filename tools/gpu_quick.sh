#!/bin/bash
# one iteration on the GPU box: parity tests, per-phase shares (A/B build), short bench
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
echo "== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
echo "== phases"; timeout 200 python tools/gpu_phases.py 2>&1 | tail -3 | tee gpurun_out/phases.json
echo "== bench"; QZ_BENCH_NOCPU=1 timeout 600 python bench.py --steps 3 > gpurun_out/bench_quick.json 2> gpurun_out/bench_err.log; python -c "
import json; d=json.load(open('gpurun_out/bench_quick.json')); print('value', d['value'], 'e2e', d['e2e']['value'], 'ratio', d['ratio'], 'ms/launch', d['roofline']['ms_per_launch'], 'frac', d['roofline']['frac'])"; tail -3 gpurun_out/bench_err.log
