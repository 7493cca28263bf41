#!/bin/bash
# A/B on one GPU box: per-piece deflate blocks (QZB200_GROUP=0) against the group kernel (QZB200_GROUP=1) at several
# geometries and hash-table sizes and in the A/B builds of qz_deflate.cu (make ab: _w32 = 32-warp launch bound,
# _pipe = software-pipelined match loop), the parity suite under the candidate default and under the pipelined
# build, the host link's duplex rates, one call's per-batch timeline.
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
echo "== nproc $(nproc)"; nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
t0=$(date +%s)
echo "== geometry points (group, hash bits, warps, buffers; 0 buffers = as many as fit)"
: > gpurun_out/geom41.jsonl
for cfg in "0 11 20 17 libqatzip.so" "1 11 24 0 libqatzip.so" "1 10 24 0 libqatzip.so" "1 10 24 18 libqatzip.so" "1 10 16 16 libqatzip.so" \
           "0 11 20 17 libqatzip_pipe.so" "1 11 24 0 libqatzip_pipe.so" "1 10 24 0 libqatzip_pipe.so" \
           "1 10 32 0 libqatzip_w32.so" "1 10 32 15 libqatzip_w32.so" "1 11 32 0 libqatzip_w32.so" "1 10 32 0 libqatzip_pipe32.so"; do
  set -- $cfg
  [ -f qatzip_b200/$5 ] || continue
  QZ_PRODUCT_SO=$PWD/qatzip_b200/$5 QZB200_GROUP=$1 QZB200_GROUP_HASH_BITS=$2 QZB200_WARPS=$3 QZB200_BUFFERS=$4 timeout 120 python tools/gpu_geom.py 2>&1 | tail -1 | sed "s/^{/{\"lib\": \"$5\", \"group\": $1, \"hb\": $2, /" | tee -a gpurun_out/geom41.jsonl
done
echo "   took $(( $(date +%s) - t0 )) s"; t0=$(date +%s)
echo "== pytest gpu, group kernel, 2^10 table"; QZB200_GROUP=1 QZB200_GROUP_HASH_BITS=10 timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_group1_hb10.log
echo "   took $(( $(date +%s) - t0 )) s"; t0=$(date +%s)
if [ -f qatzip_b200/libqatzip_pipe.so ]; then
echo "== pytest gpu, pipelined build, group kernel 2^10"; QZ_PRODUCT_SO=$PWD/qatzip_b200/libqatzip_pipe.so QZB200_GROUP=1 QZB200_GROUP_HASH_BITS=10 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_pipe_group1.log
echo "== pytest gpu, pipelined build, per-piece"; QZ_PRODUCT_SO=$PWD/qatzip_b200/libqatzip_pipe.so QZB200_GROUP=0 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_pipe_group0.log
echo "   took $(( $(date +%s) - t0 )) s"; t0=$(date +%s)
fi
echo "== host link"; timeout 120 python tools/gpu_pcie_duplex.py 2>&1 | tail -1 | tee gpurun_out/pcie_duplex.json
echo "== timeline of one host call"; timeout 120 python tools/gpu_timeline.py 2>&1 | tail -32 | tee gpurun_out/timeline.log
echo "   took $(( $(date +%s) - t0 )) s"
