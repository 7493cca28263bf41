#!/bin/bash
# window kernel: parity subset, table sizes, phases
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
TAG=${TAG:-r02c}
echo "== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
echo "== geometry sweep"
for cfg in "2 1344" "2 1024" "1 1344" "1 2048"; do set -- $cfg
  echo -n "groups $1 tent $2: "; QZB200_WINDOW_GROUPS=$1 QZB200_WINDOW_TENT=$2 timeout 120 python tools/gpu_geom.py 2>&1 | tail -1; done | tee gpurun_out/${TAG}_window_geometry.log
echo -n "per-piece kernel: "; QZB200_WINDOW=0 timeout 120 python tools/gpu_geom.py 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_window_geometry.log
echo "== phases"; timeout 200 python tools/gpu_phases.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_phases_window.json
