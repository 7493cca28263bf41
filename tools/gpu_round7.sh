#!/bin/bash
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench.json; tail -3 gpurun_out/bench_err.log
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/bench_err.log | tee gpurun_out/bench_ref.json
echo "== extra perf"; timeout 900 python tools/gpu_perf_extra.py 2>gpurun_out/extra_err.log | tee gpurun_out/extra.json; tail -3 gpurun_out/extra_err.log
echo "== ncu launch list"
QZ_BENCH_NOCPU=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --gib 1 > gpurun_out/ncu_launch_run.log 2>&1; tail -1 gpurun_out/ncu_launch_run.log | cut -c1-120
echo "== ncu full deflate (512 MiB device-resident launch)"
QZ_BENCH_NOCPU=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:qzb_deflate_pieces -s 1 -c 1 -o gpurun_out/prof_deflate_v4 -f \
   python bench.py --steps 1 --warmup 1 --gib 0.5 > gpurun_out/ncu_full_run.log 2>&1; tail -2 gpurun_out/ncu_full_run.log
