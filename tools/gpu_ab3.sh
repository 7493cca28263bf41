#!/bin/bash
# one short visit: per-phase shares of warp time in the group kernel (phase-clock build), streaming slot stores A/B
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
echo "== phases, group kernel"; timeout 200 python tools/gpu_phases.py 2>&1 | tail -1 | tee gpurun_out/phases_group.json
echo "== phases, per-piece kernel"; QZB200_GROUP=0 timeout 200 python tools/gpu_phases.py 2>&1 | tail -1 | tee gpurun_out/phases_piece.json
echo "== slot stores: plain vs st.global.cs"
for i in 1 2; do for lib in libqatzip.so libqatzip_cs.so; do
  echo -n "$lib: "; QZ_PRODUCT_SO=$PWD/qatzip_b200/$lib timeout 120 python tools/gpu_geom.py 2>&1 | tail -1
done; done | tee gpurun_out/slot_cs_ab.log
