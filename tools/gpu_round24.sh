#!/bin/bash
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
echo "== pytest gpu (parity)"; timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -3
for g in "22 13" "22 16" "24 15" "23 16" "26 14" "20 17" "25 15" "28 13" "21 17"; do set -- $g; QZB200_WARPS=$1 QZB200_BUFFERS=$2 timeout 120 python tools/gpu_geom.py 2>&1 | tail -1; done | tee gpurun_out/geom.jsonl
echo "== phases"; timeout 200 python tools/gpu_phases.py 2>&1 | tail -1
