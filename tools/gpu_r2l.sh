#!/bin/bash
# Decompress of mixed members: per-batch timeline of a host-to-host call, then the bench's inflate leg (8 GiB).
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
TAG=${TAG:-r02l}
echo "== timeline"; TL_MIB=${TL_MIB:-2048} timeout 300 python tools/gpu_infl_timeline.py 2>&1 | grep -- "--- rep" | tee gpurun_out/${TAG}_inflate_host_calls.log
echo "== inflate leg"; timeout 900 python - <<'PY' | tee gpurun_out/${TAG}_inflate_leg.json
import json, os, sys
sys.path.insert(0, os.getcwd())
from harness import qzapi as q, bench_secondary as bs
prod, cor = q.QzLib(q.PRODUCT_SO), q.Corpus()
r = bs.inflate_leg(prod, q.REF_SO, cor, 6545.0, os.cpu_count(), 2)
r.pop("cpu_baseline", None)
print(json.dumps(r))
PY
