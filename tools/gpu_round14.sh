#!/bin/bash
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
echo "== zlib diag"; timeout 240 python tools/gpu_zlib_diag.py 2>&1 | tail -40 | tee gpurun_out/zlib_diag.log
echo "== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu -x --durations=12 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
for i in 1 2; do
  for v in libqatzip.so libqatzip_ab.so; do
    echo "== bench $v #$i"; QZ_PRODUCT_SO=$PWD/qatzip_b200/$v QZ_BENCH_NOCPU=1 timeout 300 python bench.py --steps 3 2>> gpurun_out/bench_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['ms_per_launch'], d['e2e']['value'])"
  done
done
echo "== extra"; EXTRA_NOCPU=1 timeout 600 python tools/gpu_perf_extra.py > gpurun_out/extra.json 2> gpurun_out/extra_err.log; cat gpurun_out/extra.json; tail -3 gpurun_out/extra_err.log
