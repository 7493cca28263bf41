#!/bin/bash
# all GPUs of the box: host-link ceiling per N, torchrun bench at N (headline + LZ4 leg + one-process leg)
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
TAG=${TAG:-r02}
N=$(nvidia-smi -L | wc -l); echo "GPUs $N, host threads $(nproc)"
for n in 1 2 4 8; do [ $n -le $N ] || continue
  echo -n "ceiling N=$n: "; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29520 tools/gpu_pcie_ceiling.py 2>/dev/null | tail -1; done | tee gpurun_out/${TAG}_pcie_ceiling.log
echo "== torchrun bench N=$N"
QZ_BENCH_NOCPU=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/bench_err.log; tail -3 gpurun_out/bench_err.log
python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench_${N}gpu.json').read().strip().splitlines()[-1]); s=d.pop('secondary'); print(json.dumps(d)); print(json.dumps(s, indent=1))"
