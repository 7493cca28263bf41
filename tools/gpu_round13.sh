#!/bin/bash
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
echo "== zlib diag"; timeout 600 python tools/gpu_zlib_diag.py 2>&1 | tail -60 | tee gpurun_out/zlib_diag.log
