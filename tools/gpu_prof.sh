#!/bin/bash
# quick: bench (1 GiB) + one full ncu capture of the deflate kernel on a 64 MiB host batch
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "all_sizes or stream or errors or roundtrip" 2>&1 | tail -3
QZ_BENCH_NOCPU=1 timeout 300 python bench.py --steps 3 --warmup 2 --gib 2 2>gpurun_out/bench_err.log | tee gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_err.log
QZ_BENCH_NOCPU=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:qzb_deflate_pieces -s 3 -c 1 -o gpurun_out/prof_deflate -f \
   python bench.py --steps 1 --warmup 1 --gib 0.5 > gpurun_out/ncu_full_run.log 2>&1; tail -2 gpurun_out/ncu_full_run.log
