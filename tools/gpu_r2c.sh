#!/bin/bash
# window kernel shapes: warps per window x groups x units x table entries; phases of the default
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
TAG=${TAG:-r02c}
echo "== pytest gpu (quick subset)"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "all_sizes or ratio or roundtrip" 2>&1 | tail -3
echo "== geometry sweep"
for cfg in "16 2 2 1344" "16 2 2 1024" "16 2 2 768" "16 1 1 1344" "8 4 2 2048" "8 3 2 2048" "8 4 2 1024"; do set -- $cfg
  echo -n "warps/window $1 groups $2 units $3 tent $4: "; QZB200_WINDOW_WARPS=$1 QZB200_WINDOW_GROUPS=$2 QZB200_WINDOW_UNITS=$3 QZB200_WINDOW_TENT=$4 timeout 120 python tools/gpu_geom.py 2>&1 | tail -1; done | tee gpurun_out/${TAG}_window_geometry.log
echo -n "per-piece kernel: "; QZB200_WINDOW=0 timeout 120 python tools/gpu_geom.py 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_window_geometry.log
echo "== phases"; timeout 200 python tools/gpu_phases.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_phases_window.json
QZB200_WINDOW=0 timeout 200 python tools/gpu_phases.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_phases_pieces.json
