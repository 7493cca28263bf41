#!/bin/bash
# deflate window kernel: parity suite, timing at level 1 and 9, phases
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
TAG=${TAG:-r02c}
echo "== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
echo "== deflate"
for lvl in 1 9; do echo -n "level $lvl: "; GEOM_LEVEL=$lvl timeout 120 python tools/gpu_geom.py 2>&1 | tail -1; done | tee gpurun_out/${TAG}_window_levels.log
echo "== phases"; timeout 200 python tools/gpu_phases.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_phases_window.json
