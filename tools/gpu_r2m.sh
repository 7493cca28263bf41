#!/bin/bash
# LZ4 geometry by level: ratio gates, level test, throughput of the LZ4 leg at level 1.
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
TAG=${TAG:-r02m}
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "ratio or levels or lz4 or LZ4" 2>&1 | grep -E "deflate |LZ4 |levels|passed|failed|Error" | tee gpurun_out/${TAG}_ratio_gates.log
echo "== lz4 leg"; timeout 600 python - <<'PY' | tee gpurun_out/${TAG}_lz4_leg.json
import json, os, sys
sys.path.insert(0, os.getcwd())
from harness import qzapi as q, bench_secondary as bs
prod, cor = q.QzLib(q.PRODUCT_SO), q.Corpus()
ident = lambda x: x
r = bs.lz4_leg(prod, None, cor, 6545.0, os.cpu_count(), 3, 0, 1, lambda: None, ident, ident)
print(json.dumps(r))
PY
