#!/bin/bash
# inflate: parity suite, then decoders per warp over our members and over reference-made mixed members
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
TAG=${TAG:-r02e}
echo "== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
for d in 4 8 2 1; do echo "== decoders per warp $d"; QZB200_INFLATE_DPW=$d INFL_MIB=2048 INFL_REF_MIB=2048 INFL_REPS=3 timeout 600 python tools/gpu_inflate_bench.py 2>&1 | tail -1; done | tee gpurun_out/${TAG}_inflate_dpw.log
