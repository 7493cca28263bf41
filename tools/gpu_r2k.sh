#!/bin/bash
# Rounds inflate kernel: correctness under both slot counts, then throughput.
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
TAG=${TAG:-r02k}
for d in 31 16; do
echo "== pytest gpu (decode side), QZB200_INFLATE_DPW=$d"; QZB200_INFLATE_DPW=$d timeout 600 python -m pytest tests -q -m gpu -x -k "decomp or infl or golden or stream or roundtrip or corrupt or error" 2>&1 | tail -4 | tee -a gpurun_out/${TAG}_pytest.log
done
for so in ${SOS:-libqatzip.so}; do for d in ${DPWS:-31 16 1}; do
  echo "== $so QZB200_INFLATE_DPW=$d"
  QZ_PRODUCT_SO=$PWD/qatzip_b200/$so QZB200_INFLATE_DPW=$d INFL_MIB=2048 INFL_REF_MIB=2048 INFL_REPS=3 timeout 300 python tools/gpu_inflate_bench.py 2>&1 | tail -1 | cut -c1-420
done; done | tee gpurun_out/${TAG}_inflate_rounds.log
