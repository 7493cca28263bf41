#!/bin/bash
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
INFL_MIB=512 INFL_REPS=2 INFL_CASES=ours timeout 600 ncu --set full --clock-control none --import-source on -k regex:qzb_inflate_kernel -s 1 -c 1 -o gpurun_out/prof_inflate -f python tools/gpu_inflate_bench.py > gpurun_out/ncu_inflate.log 2>&1; tail -2 gpurun_out/ncu_inflate.log
