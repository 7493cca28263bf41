#!/bin/bash
# A/B on one GPU box: per-piece deflate blocks (QZB200_GROUP=0) against the group kernel (QZB200_GROUP=1):
# parity suite under both, geometry points, headline bench under both.  Everything lands in gpurun_out/.
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
echo "== nproc $(nproc)"; nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
t0=$(date +%s)
echo "== pytest gpu, group kernel"; QZB200_GROUP=1 timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_group1.log
echo "   took $(( $(date +%s) - t0 )) s"; t0=$(date +%s)
echo "== pytest gpu, per-piece kernel"; QZB200_GROUP=0 timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_group0.log
echo "   took $(( $(date +%s) - t0 )) s"; t0=$(date +%s)
echo "== geometry points"
: > gpurun_out/geom40.jsonl
for cfg in "0 20 17" "0 24 15" "1 24 15" "1 24 14" "1 16 16" "1 16 13" "1 24 12"; do
  set -- $cfg
  QZB200_GROUP=$1 QZB200_WARPS=$2 QZB200_BUFFERS=$3 timeout 120 python tools/gpu_geom.py 2>&1 | tail -1 | sed "s/^{/{\"group\": $1, /" | tee -a gpurun_out/geom40.jsonl
done
echo "   took $(( $(date +%s) - t0 )) s"; t0=$(date +%s)
echo "== bench group=1"; QZB200_GROUP=1 timeout 600 python bench.py > gpurun_out/bench_group1.json 2> gpurun_out/bench_group1_err.log; cat gpurun_out/bench_group1.json; tail -2 gpurun_out/bench_group1_err.log
echo "   took $(( $(date +%s) - t0 )) s"; t0=$(date +%s)
echo "== bench group=0"; QZ_BENCH_NOCPU=1 QZB200_GROUP=0 timeout 600 python bench.py > gpurun_out/bench_group0.json 2> gpurun_out/bench_group0_err.log; cat gpurun_out/bench_group0.json; tail -2 gpurun_out/bench_group0_err.log
echo "   took $(( $(date +%s) - t0 )) s"
