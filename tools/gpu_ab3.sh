#!/bin/bash
# one short visit: builds and group shapes interleaved on the geometry tool, parity suite under both shapes, phase shares
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
: > gpurun_out/ab3.log
for i in 1 2; do
  for cfg in "libqatzip_old.so 8 0" "libqatzip.so 8 0" "libqatzip.so 4 0" "libqatzip.so 4 17" "libqatzip.so 8 17"; do
    set -- $cfg; [ -f qatzip_b200/$1 ] || continue
    echo -n "$1 group_warps=$2 bufs=$3: " | tee -a gpurun_out/ab3.log
    QZ_PRODUCT_SO=$PWD/qatzip_b200/$1 QZB200_GROUP_WARPS=$2 QZB200_BUFFERS=$3 timeout 120 python tools/gpu_geom.py 2>&1 | tail -1 | tee -a gpurun_out/ab3.log
  done
done
echo "== pytest gpu, 4 warps per group"; QZB200_GROUP_WARPS=4 timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
echo "== pytest gpu, 8 warps per group"; QZB200_GROUP_WARPS=8 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -3
echo "== phases, 4 warps per group"; QZB200_GROUP_WARPS=4 timeout 200 python tools/gpu_phases.py 2>&1 | tail -1 | tee gpurun_out/phases_group4.json
echo "== phases, 8 warps per group"; QZB200_GROUP_WARPS=8 timeout 200 python tools/gpu_phases.py 2>&1 | tail -1 | tee gpurun_out/phases_group8.json
