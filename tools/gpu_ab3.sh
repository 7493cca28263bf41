#!/bin/bash
# one short visit: two builds of the library interleaved on the geometry tool, parity subset on the new one, phase shares
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
for i in 1 2 3; do for lib in ${AB_LIBS:-libqatzip_old.so libqatzip.so}; do
  [ -f qatzip_b200/$lib ] || continue
  echo -n "$lib: "; QZ_PRODUCT_SO=$PWD/qatzip_b200/$lib timeout 120 python tools/gpu_geom.py 2>&1 | tail -1
done; done | tee gpurun_out/ab3.log
echo "== pytest gpu (compress-side subset)"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "ours_to_oracle or ratio or round_trip or crc or static or stream_compress" 2>&1 | tail -3
echo "== phases, group kernel"; timeout 200 python tools/gpu_phases.py 2>&1 | tail -1 | tee gpurun_out/phases_group.json
echo "== phases, per-piece kernel"; QZB200_GROUP=0 timeout 200 python tools/gpu_phases.py 2>&1 | tail -1 | tee gpurun_out/phases_piece.json
