#!/bin/bash
# A/B on one GPU box: per-piece deflate blocks (QZB200_GROUP=0) against the group kernel (QZB200_GROUP=1) at several
# geometries and hash-table sizes, the parity suite under the candidate default, the host link's duplex rates.
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
echo "== nproc $(nproc)"; nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
t0=$(date +%s)
echo "== pytest gpu, group kernel, 2^10 table"; QZB200_GROUP=1 QZB200_GROUP_HASH_BITS=10 timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_group1_hb10.log
echo "   took $(( $(date +%s) - t0 )) s"; t0=$(date +%s)
echo "== geometry points (group, hash bits, warps, buffers; 0 buffers = as many as fit)"
: > gpurun_out/geom41.jsonl
for cfg in "0 11 20 17 libqatzip.so" "1 11 24 0 libqatzip.so" "1 10 24 0 libqatzip.so" "1 10 24 18 libqatzip.so" "1 10 24 16 libqatzip.so" "1 10 16 16 libqatzip.so" \
           "1 10 32 0 libqatzip_ab.so" "1 10 32 15 libqatzip_ab.so" "1 11 32 0 libqatzip_ab.so" "0 11 20 17 libqatzip_ab.so"; do
  set -- $cfg
  QZ_PRODUCT_SO=$PWD/qatzip_b200/$5 QZB200_GROUP=$1 QZB200_GROUP_HASH_BITS=$2 QZB200_WARPS=$3 QZB200_BUFFERS=$4 timeout 120 python tools/gpu_geom.py 2>&1 | tail -1 | sed "s/^{/{\"lib\": \"$5\", \"group\": $1, \"hb\": $2, /" | tee -a gpurun_out/geom41.jsonl
done
echo "   took $(( $(date +%s) - t0 )) s"; t0=$(date +%s)
echo "== host link"; timeout 120 python tools/gpu_pcie_duplex.py 2>&1 | tail -1 | tee gpurun_out/pcie_duplex.json
echo "   took $(( $(date +%s) - t0 )) s"
