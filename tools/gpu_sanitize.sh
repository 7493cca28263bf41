#!/bin/bash
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
echo "== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
echo "== racecheck on a small compress (shared-memory hazards)"
DIAG_QUICK=1 DIAG_FMTS=2 timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/gpu_diag.py > gpurun_out/racecheck.log 2>&1; grep -E "RACECHECK SUMMARY|hazard|ERROR SUMMARY|FAIL" gpurun_out/racecheck.log | head -8
echo "== memcheck lz4 + deflate + inflate"
DIAG_QUICK=1 DIAG_FMTS=2,4,0 timeout 900 compute-sanitizer --tool memcheck python tools/gpu_diag.py > gpurun_out/memcheck.log 2>&1; grep -E "ERROR SUMMARY|Invalid|FAIL" gpurun_out/memcheck.log | head -5
