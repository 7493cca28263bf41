#!/bin/bash
# window kernel: parity suite, table sizes, phases
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
TAG=${TAG:-r02c}
echo "== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
echo "== table sizes"
for tent in 0 2048 1344; do echo -n "tent $tent: "; QZB200_WINDOW_TENT=$tent timeout 120 python tools/gpu_geom.py 2>&1 | tail -1; done | tee gpurun_out/${TAG}_window_geometry.log
echo -n "per-piece kernel: "; QZB200_WINDOW=0 timeout 120 python tools/gpu_geom.py 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_window_geometry.log
echo "== phases"; timeout 200 python tools/gpu_phases.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_phases_window.json
