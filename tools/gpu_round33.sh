#!/bin/bash
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
INFL_MIB=2048 INFL_REF_MIB=2048 INFL_REPS=3 INFL_CASES=ref timeout 900 python tools/gpu_inflate_bench.py 2>&1 | tail -3 | tee gpurun_out/inflate_bench_big.json
