#!/bin/bash
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
echo "== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench.json; tail -3 gpurun_out/bench_err.log
echo "== memcheck (small)"; DIAG_QUICK=1 DIAG_FMTS=2,3,4 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/gpu_diag.py > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|Invalid|FAIL|EXC" gpurun_out/memcheck.log | head -10
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  echo "== bench 2 GPUs"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 2>gpurun_out/bench2_err.log | tee gpurun_out/bench_n2.json; tail -3 gpurun_out/bench2_err.log
fi
