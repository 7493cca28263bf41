#!/bin/bash
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
echo "== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench.json; tail -3 gpurun_out/bench_err.log
echo "== warps sweep (2 GiB)"
for w in 16 20 24; do echo "-- warps=$w"; QZB200_WARPS=$w QZ_BENCH_NOCPU=1 timeout 300 python bench.py --steps 3 --warmup 2 --gib 2 2>>gpurun_out/bench_err.log | tee -a gpurun_out/sweep.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['roofline']['ms_per_launch'])"; done
echo "== extra perf"; timeout 900 python tools/gpu_perf_extra.py 2>gpurun_out/extra_err.log | tee gpurun_out/extra.json; tail -3 gpurun_out/extra_err.log
echo "== ncu full (512 MiB device-resident launch)"
QZ_BENCH_NOCPU=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:qzb_deflate_pieces -s 1 -c 1 -o gpurun_out/prof_deflate_v2 -f \
   python bench.py --steps 1 --warmup 1 --gib 0.5 > gpurun_out/ncu_full_run.log 2>&1; tail -2 gpurun_out/ncu_full_run.log
echo "== ncu launch list"
QZ_BENCH_NOCPU=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --gib 0.5 > gpurun_out/ncu_launch_run.log 2>&1; tail -1 gpurun_out/ncu_launch_run.log | cut -c1-200
