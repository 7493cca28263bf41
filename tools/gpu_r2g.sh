#!/bin/bash
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
TAG=${TAG:-r02g}
echo "== pytest gpu"; timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -12 | tee gpurun_out/${TAG}_pytest_gpu.log
echo "== stream"; timeout 300 python - <<'PY'
import json, os, sys
sys.path.insert(0, os.getcwd())
from harness import qzapi as q, bench_secondary as bs
prod, cor = q.QzLib(q.PRODUCT_SO), q.Corpus()
os.environ["QZ_BENCH_STREAM_GIB"] = "1"
print(json.dumps(bs.stream_leg(prod, q.REF_SO if os.path.exists(q.REF_SO) else None, cor, 6545.0, os.cpu_count())))
for kb in ("4096", "16384", "32768"):
    os.environ["QZB200_STREAM_BATCH_KB"] = kb
    r = bs.stream_leg(prod, None, cor, 6545.0, os.cpu_count()); print(kb, r["value"], r["stream_crc_matches_zlib"])
PY
