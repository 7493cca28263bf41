#!/bin/bash
# First contact of the experimental matcher / coder kernel (qz_deflate_split.cuh) with a GPU.  Before the visit, here:
#   make -C qatzip_b200/csrc ab ABFLAGS=-DQZ_SPLIT_KERNEL ABNAME=split          (-> qatzip_b200/libqatzip_split.so)
# On the box: parity suite through the split build with QZB200_GROUP=2, then matcher/team geometries against the group kernel.
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
SPLIT=$PWD/qatzip_b200/libqatzip_split.so
[ -f $SPLIT ] || { echo "build libqatzip_split.so first"; exit 1; }
echo "== pytest gpu through the split kernel"; QZ_PRODUCT_SO=$SPLIT QZB200_GROUP=2 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_split.log
echo "== geometry points"
: > gpurun_out/split_ab.jsonl
echo -n "group kernel: " | tee -a gpurun_out/split_ab.jsonl; QZB200_GROUP=1 timeout 120 python tools/gpu_geom.py 2>&1 | tail -1 | tee -a gpurun_out/split_ab.jsonl
for cfg in "19 3" "18 3" "20 3" "16 4" "20 2" "17 3"; do
  set -- $cfg
  echo -n "split matchers=$1 teams=$2: " | tee -a gpurun_out/split_ab.jsonl
  QZ_PRODUCT_SO=$SPLIT QZB200_GROUP=2 QZB200_SPLIT_MATCHERS=$1 QZB200_SPLIT_TEAMS=$2 timeout 120 python tools/gpu_geom.py 2>&1 | tail -1 | tee -a gpurun_out/split_ab.jsonl
done
