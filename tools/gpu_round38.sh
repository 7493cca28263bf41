#!/bin/bash
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
QZ_BENCH_NOCPU=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:qzb_deflate_pieces -s 1 -c 1 -o gpurun_out/prof_deflate -f \
   python bench.py --steps 1 --warmup 1 --gib 0.5 > gpurun_out/ncu_full_run.log 2>&1; tail -1 gpurun_out/ncu_full_run.log
echo "== inflate bench"; INFL_MIB=2048 INFL_REF_MIB=2048 timeout 900 python tools/gpu_inflate_bench.py 2>&1 | tail -1 | tee gpurun_out/inflate_bench.json
