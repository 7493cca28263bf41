#!/bin/bash
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
echo "== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu -x --durations=8 2>&1 | tail -16 | tee gpurun_out/pytest_gpu.log
echo "== phases"; timeout 200 python tools/gpu_phases.py 2>&1 | tail -3 | tee gpurun_out/phases.json
echo "== bench"; timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench_err.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench_err.log
echo "== extra"; EXTRA_NOCPU=1 timeout 600 python tools/gpu_perf_extra.py > gpurun_out/extra.json 2> gpurun_out/extra_err.log; cat gpurun_out/extra.json; tail -3 gpurun_out/extra_err.log
