#!/bin/bash
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
TAG=${TAG:-r02h}
echo "== pytest gpu"; timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest_gpu.log
echo "== ratio lines"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "ratio" 2>&1 | grep -E "deflate |LZ4 " | tee gpurun_out/${TAG}_ratio_gates.log
echo "== inflate"; QZB200_INFLATE_DPW=1 INFL_MIB=2048 INFL_REF_MIB=2048 INFL_REPS=3 timeout 600 python tools/gpu_inflate_bench.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_inflate_bench.json
echo "== inflate leg (8 GiB, large calls)"; timeout 600 python - <<'PY'
import json, os, sys
sys.path.insert(0, os.getcwd())
from harness import qzapi as q, bench_secondary as bs
prod, cor = q.QzLib(q.PRODUCT_SO), q.Corpus()
print(json.dumps(bs.inflate_leg(prod, q.REF_SO, cor, 6545.0, os.cpu_count(), 2)))
PY
