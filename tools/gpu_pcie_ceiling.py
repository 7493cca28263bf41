"""The box's end-to-end ceiling for N processes: every rank of a torchrun launch copies its share of the headline workload's
bytes (512 MiB in, 0.41 x that out per call) H2D and D2H at once from and to its own pinned pages, all ranks between two
barriers.  The aggregate input rate is what `e2e` of bench.py could reach if the kernels cost nothing."""
import json
import os
import time

import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_in, n_out, reps = 512 << 20, 212 << 20, 8
h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n_in, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n_out, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def once():
    with torch.cuda.stream(s1):
        for _ in range(reps):
            d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        for _ in range(reps):
            h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()


once()
if world > 1:
    dist.barrier()
torch.cuda.synchronize(); t0 = time.perf_counter()
once()
if world > 1:
    dist.barrier()
dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
if int(os.environ.get("RANK", "0")) == 0:
    print(json.dumps({"n_gpus": world, "aggregate_input_GBps_with_output_flowing_back": round(world * reps * n_in / dt.item() / 1e9, 2),
                      "per_gpu": round(reps * n_in / dt.item() / 1e9, 2), "bytes": {"h2d_per_rank": reps * n_in, "d2h_per_rank": reps * n_out}}))
if world > 1:
    dist.destroy_process_group()
