#!/bin/bash
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
echo "== pytest gpu (parity subset)"; timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -3
for g in "22 13" "20 14" "18 15" "23 12" "21 13" "16 16" "24 11" "19 14"; do set -- $g; QZB200_WARPS=$1 QZB200_BUFFERS=$2 timeout 120 python tools/gpu_geom.py 2>&1 | tail -1; done | tee gpurun_out/geom.jsonl
