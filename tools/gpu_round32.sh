#!/bin/bash
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
echo "== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
echo "== extra"; EXTRA_NOCPU=1 EXTRA_STREAM_MIB=256 timeout 900 python tools/gpu_perf_extra.py > gpurun_out/extra.json 2> gpurun_out/extra_err.log; cat gpurun_out/extra.json; tail -3 gpurun_out/extra_err.log
