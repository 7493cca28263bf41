#!/bin/bash
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
TAG=${TAG:-r02e}
for so in libqatzip.so libqatzip_lut97.so; do for d in 1 2; do echo "== $so decoders per warp $d"; QZ_PRODUCT_SO=$PWD/qatzip_b200/$so QZB200_INFLATE_DPW=$d INFL_MIB=2048 INFL_REF_MIB=2048 INFL_REPS=3 timeout 600 python tools/gpu_inflate_bench.py 2>&1 | tail -1; done; done | tee gpurun_out/${TAG}_inflate_dpw.log
