#!/bin/bash
# two GPUs: one-process multi-device tests, then a short torchrun bench (secondary legs at small sizes)
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
TAG=${TAG:-r02f}
nvidia-smi -L
echo "== pytest multi"; timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/${TAG}_pytest_multi.log
echo "== torchrun bench 2 GPUs"
QZ_BENCH_LZ4_GIB=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 --gib 2 > gpurun_out/${TAG}_bench_2gpu.json 2> gpurun_out/bench_err.log; tail -3 gpurun_out/bench_err.log
python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench_2gpu.json').read().strip().splitlines()[-1]); print('value', d['value'], 'e2e', d['e2e']['value']); print(json.dumps(d['secondary'], indent=1)[:3000])"
