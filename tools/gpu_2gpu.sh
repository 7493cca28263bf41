#!/bin/bash
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
N=${1:-2}
QZ_BENCH_NOCPU=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu_err.log; cat gpurun_out/bench_${N}gpu.json; tail -3 gpurun_out/bench_${N}gpu_err.log
