#!/bin/bash
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
for v in ${AB_LIBS:-libqatzip.so libqatzip_old.so}; do
  for g in "22 13" "20 14" "21 13" "20 13"; do set -- $g
    echo -n "$v: "; QZB200_WARPS=$1 QZB200_BUFFERS=$2 QZ_PRODUCT_SO=$PWD/qatzip_b200/$v timeout 120 python tools/gpu_geom.py 2>&1 | tail -1
  done
done | tee gpurun_out/ab2.log
