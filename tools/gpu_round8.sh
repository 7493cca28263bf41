#!/bin/bash
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
echo "== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
echo "== pytest gpu (warp-per-member inflate forced)"; QZB200_INFLATE_LANE_MIN=1000000000 timeout 900 python -m pytest tests -q -m gpu -k "oracle_to_ours or mixed or errors or chunk_sizes or large" 2>&1 | tail -3
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench.json; tail -3 gpurun_out/bench_err.log
echo "== extra perf"; timeout 900 python tools/gpu_perf_extra.py 2>gpurun_out/extra_err.log | tee gpurun_out/extra.json; tail -3 gpurun_out/extra_err.log
echo "== ncu dram traffic of the deflate kernel (512 MiB launch)"
QZ_BENCH_NOCPU=1 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:qzb_deflate_pieces -s 1 -c 1 --csv --log-file gpurun_out/traffic.csv python bench.py --steps 1 --warmup 1 --gib 0.5 > gpurun_out/ncu_traffic_run.log 2>&1; grep -E "dram__|gpu__time" gpurun_out/traffic.csv | cut -d, -f12- | head
