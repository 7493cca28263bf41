#!/bin/bash
# Stream compress with one and two workers, several staging sizes; then the stream tests.
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
TAG=${TAG:-r02i}
echo "== stream tests"; timeout 900 python -m pytest tests -q -m gpu -x -k "stream" 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest_stream.log
export QZ_BENCH_STREAM_GIB=${QZ_BENCH_STREAM_GIB:-4}
for W in 1 2; do for KB in 4096 8192 4096 8192; do
echo "== workers $W batch $KB KiB"
QZB200_STREAM_WORKERS=$W QZB200_STREAM_BATCH_KB=$KB timeout 600 python - <<'PY' 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_stream_sweep.log
import json, os, sys
sys.path.insert(0, os.getcwd())
from harness import qzapi as q, bench_secondary as bs
prod, cor = q.QzLib(q.PRODUCT_SO), q.Corpus()
r = bs.stream_leg(prod, None, cor, 6545.0, os.cpu_count())
print(json.dumps({"ceiling": r.get("host_copy_ceiling",{}).get("value"), "workers": os.environ["QZB200_STREAM_WORKERS"], "batch_kb": os.environ["QZB200_STREAM_BATCH_KB"], "value": r["value"], "ratio": r["ratio"], "crc_ok": r["stream_crc_matches_zlib"]}))
PY
done; done
