#!/bin/bash
# one full ncu capture of the deflate kernel in use (TAG names the report)
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
TAG=${TAG:-r02}
QZ_BENCH_NOCPU=1 timeout 900 ncu --set full --clock-control none --import-source on -k "regex:qzb_deflate_(window|pieces)" -s 1 -c 1 -o gpurun_out/${TAG}_prof_deflate -f \
   python bench.py --steps 1 --warmup 1 --gib 0.5 > gpurun_out/ncu_full_run.log 2>&1; tail -1 gpurun_out/ncu_full_run.log
