#!/bin/bash
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
echo "== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench.json; tail -3 gpurun_out/bench_err.log
echo "== extra perf"; timeout 900 python tools/gpu_perf_extra.py 2>gpurun_out/extra_err.log | tee gpurun_out/extra.json; tail -3 gpurun_out/extra_err.log
echo "== ncu full deflate (512 MiB device-resident launch)"
QZ_BENCH_NOCPU=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:qzb_deflate_pieces -s 1 -c 1 -o gpurun_out/prof_deflate_v3 -f \
   python bench.py --steps 1 --warmup 1 --gib 0.5 > gpurun_out/ncu_full_run.log 2>&1; tail -2 gpurun_out/ncu_full_run.log
echo "== ncu full inflate"
EXTRA_MIB=512 EXTRA_STREAM_MIB=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:qzb_inflate -s 2 -c 1 -o gpurun_out/prof_inflate_v2 -f \
   python tools/gpu_perf_extra.py > gpurun_out/ncu_infl_run.log 2>&1; tail -2 gpurun_out/ncu_infl_run.log
