#!/bin/bash
# one full ncu capture of the inflate kernel selected by QZB200_INFLATE_DPW (TAG names the report)
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
TAG=${TAG:-r02}
INFL_MIB=${INFL_MIB:-512} INFL_CASES=${INFL_CASES:-ours} INFL_REF_MIB=${INFL_MIB:-512} INFL_REPS=2 timeout 900 ncu --set full --clock-control none --import-source on -k "regex:qzb_inflate" -s 1 -c 1 -o gpurun_out/${TAG}_prof_inflate -f \
   python tools/gpu_inflate_bench.py > gpurun_out/ncu_inflate_run.log 2>&1; tail -2 gpurun_out/ncu_inflate_run.log | cut -c1-300
