#!/bin/bash
# phases (A/B build with -DQZ_PHASE_CLOCKS) + one full ncu capture of the deflate kernel in use
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
TAG=${TAG:-r02b}
echo "== phases"; timeout 200 python tools/gpu_phases.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_phases_window.json
QZB200_WINDOW=0 timeout 200 python tools/gpu_phases.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_phases_pieces.json
echo "== ncu full: deflate kernel"
QZ_BENCH_NOCPU=1 timeout 900 ncu --set full --clock-control none --import-source on -k "regex:qzb_deflate_(window|pieces)" -s 1 -c 1 -o gpurun_out/${TAG}_prof_deflate -f \
   python bench.py --steps 1 --warmup 1 --gib 0.5 > gpurun_out/ncu_full_run.log 2>&1; tail -1 gpurun_out/ncu_full_run.log
