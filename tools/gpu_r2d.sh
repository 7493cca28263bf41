#!/bin/bash
# parity suite; deflate table sizes + phases; LZ4 window kernel shapes; extra (inflate, lz4, zlib) numbers
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
TAG=${TAG:-r02d}
echo "== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest_gpu.log
echo "== deflate"
for tent in 0 2048; do echo -n "tent $tent: "; QZB200_WINDOW_TENT=$tent timeout 120 python tools/gpu_geom.py 2>&1 | tail -1; done | tee gpurun_out/${TAG}_window_geometry.log
echo "== phases"; timeout 200 python tools/gpu_phases.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_phases_window.json
echo "== extra (lz4 12 warps)"; EXTRA_NOCPU=1 EXTRA_STREAM_MIB=0 timeout 600 python tools/gpu_perf_extra.py > gpurun_out/${TAG}_extra.json 2> gpurun_out/extra_err.log; cat gpurun_out/${TAG}_extra.json; tail -3 gpurun_out/extra_err.log
echo "== extra (lz4 16 warps)"; QZB200_LZ4_WARPS=16 EXTRA_NOCPU=1 EXTRA_STREAM_MIB=0 EXTRA_ONLY=lz4 timeout 600 python tools/gpu_perf_extra.py 2>> gpurun_out/extra_err.log | tee gpurun_out/${TAG}_extra_lz4_16.json; tail -3 gpurun_out/extra_err.log
