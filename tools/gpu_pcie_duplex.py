"""What the host link gives on this box: pinned H2D alone, D2H alone, and both at once on two streams, in the byte
proportion of the headline workload (512 MiB in, about 0.43 x that out).  The end-to-end figure of bench.py is bounded
by the H2D rate under concurrent D2H traffic, not by the one-directional rate."""
import json
import torch

n_in, n_out = 512 << 20, 220 << 20
h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n_in, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n_out, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=8):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    torch.cuda.synchronize()
    if h2d:
        with torch.cuda.stream(s1):
            ev[0].record()
            for _ in range(reps):
                d_in.copy_(h_in, non_blocking=True)
            ev[1].record()
    if d2h:
        with torch.cuda.stream(s2):
            ev[2].record()
            for _ in range(reps):
                h_out.copy_(d_out, non_blocking=True)
            ev[3].record()
    torch.cuda.synchronize()
    r = {}
    if h2d:
        r["h2d_GBps"] = round(reps * n_in / ev[0].elapsed_time(ev[1]) / 1e6, 2)
    if d2h:
        r["d2h_GBps"] = round(reps * n_out / ev[2].elapsed_time(ev[3]) / 1e6, 2)
    return r


run(True, True, 2)
print(json.dumps({"h2d_alone": run(True, False), "d2h_alone": run(False, True), "both": run(True, True),
                  "note": "both: the D2H stream finishes first (fewer bytes); its rate and the H2D rate overlap for that part only"}))
