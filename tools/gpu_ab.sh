#!/bin/bash
# A/B on one box: same geometry tool against several builds of the library, interleaved
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
for i in 1 2 3; do
  for v in ${AB_LIBS:-libqatzip.so libqatzip_old.so}; do
    echo -n "$v: "; QZ_PRODUCT_SO=$PWD/qatzip_b200/$v timeout 120 python tools/gpu_geom.py 2>&1 | tail -1
  done
done | tee gpurun_out/ab.log
