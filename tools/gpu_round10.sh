#!/bin/bash
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
echo "== pytest gpu with lane inflate"; QZB200_INFLATE_LANE_MIN=2 timeout 900 python -m pytest tests -q -m gpu -x -k "not cli" 2>&1 | tail -4
echo "== extra perf (lanes on)"; QZB200_INFLATE_LANE_MIN=64 EXTRA_NOCPU=1 EXTRA_STREAM_MIB=2 timeout 900 python tools/gpu_perf_extra.py 2>gpurun_out/extra_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['gzip_ext_decompress'])"; tail -3 gpurun_out/extra_err.log
