#!/bin/bash
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
echo "== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
timeout 600 python tools/gpu_inflate_bench.py 2>&1 | tail -3 | tee gpurun_out/inflate_bench.json
