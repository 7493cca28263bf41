#!/bin/bash
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
echo "== pytest gpu with lane inflate"; QZB200_INFLATE_LANE_MIN=2 timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -6
echo "== extra perf (lanes on)"; QZB200_INFLATE_LANE_MIN=64 EXTRA_NOCPU=1 EXTRA_STREAM_MIB=8 timeout 900 python tools/gpu_perf_extra.py 2>gpurun_out/extra_err.log | tee gpurun_out/extra_lanes.json; tail -3 gpurun_out/extra_err.log
echo "== extra perf (lanes off)"; EXTRA_NOCPU=1 EXTRA_STREAM_MIB=8 timeout 900 python tools/gpu_perf_extra.py 2>gpurun_out/extra_err.log | tee gpurun_out/extra_warp.json; tail -3 gpurun_out/extra_err.log
