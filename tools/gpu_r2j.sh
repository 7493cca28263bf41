#!/bin/bash
# Inflate after the token-loop rewrite: correctness, then builds x decoders per warp.
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
TAG=${TAG:-r02j}
echo "== pytest gpu (decode side)"; timeout 900 python -m pytest tests -q -m gpu -x -k "decomp or infl or golden or stream or roundtrip" 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest.log
for so in ${SOS:-libqatzip.so libqatzip_c4.so libqatzip_lut97.so libqatzip_lut97c4.so}; do for d in ${DPWS:-1 2}; do
  echo "== $so decoders per warp $d"
  QZ_PRODUCT_SO=$PWD/qatzip_b200/$so QZB200_INFLATE_DPW=$d INFL_MIB=2048 INFL_REF_MIB=2048 INFL_REPS=3 timeout 600 python tools/gpu_inflate_bench.py 2>&1 | tail -1 | cut -c1-420
done; done | tee gpurun_out/${TAG}_inflate_builds.log
