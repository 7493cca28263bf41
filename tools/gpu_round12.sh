#!/bin/bash
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
echo "== pytest gpu"; timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
echo "== bench (default)"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench_err.log; cat gpurun_out/bench.json
echo "== bench old batching (first 16 MiB, no taper)"; QZ_BENCH_NOCPU=1 QZB200_FIRST_MB=16 QZB200_TAPER=0 timeout 600 python bench.py --steps 3 > gpurun_out/bench_oldbatch.json 2>> gpurun_out/bench_err.log; cat gpurun_out/bench_oldbatch.json
echo "== bench first 8 MiB + taper"; QZ_BENCH_NOCPU=1 QZB200_FIRST_MB=8 timeout 600 python bench.py --steps 3 > gpurun_out/bench_f8.json 2>> gpurun_out/bench_err.log; cat gpurun_out/bench_f8.json
echo "== bench 128 MiB batches"; QZ_BENCH_NOCPU=1 QZB200_BATCH_MB=128 timeout 600 python bench.py --steps 3 > gpurun_out/bench_b128.json 2>> gpurun_out/bench_err.log; cat gpurun_out/bench_b128.json
echo "== extra"; EXTRA_NOCPU=1 timeout 900 python tools/gpu_perf_extra.py > gpurun_out/extra.json 2> gpurun_out/extra_err.log; cat gpurun_out/extra.json; tail -3 gpurun_out/extra_err.log
