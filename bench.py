#!/usr/bin/env python
"""bench.py -- the headline measurement (BASELINE.json: qzCompress GB/s of input at 64 KiB chunks).

A "step" is one pass of the hot path over the whole per-GPU workload:
    qzCompress, QZ_DEFLATE_GZIP_EXT, level 1, hw_buff_sz 64 KiB, 4 GiB SILESIA-LIKE synthetic,
    issued as 8 calls of 512 MiB (the API's lengths are 32-bit), BASELINE.json configs[1].
Every rank owns one GPU and its own 4 GiB shard (weak scaling, no data-path collective).

  value      input bytes / second with the input already resident in HBM (device-resident entry
             point qzb200CompressDevice -> same kernels), all ranks summed, max-over-ranks time.
  e2e        the same pass through the reference-facing C ABI: qzCompress() with HOST buffers
             from qzMalloc(PINNED), host->device and device->host copies inside the timed region.
             One submitting thread by default, every call synchronous.  --e2e-threads T issues the calls from
             T threads with one session each (the pattern of the reference's own benchmark, test/main.c:2175-2202);
             on B200 that measured the same (profiles/r01_v7_bench_2threads.json): the host link under H2D + D2H
             traffic is what bounds a call (profiles/r01_v7_pcie_duplex.json, r01_v7_e2e_timeline.log).
  roofline   (bytes in + bytes out) of one deflate-kernel launch / its CUDA-event duration,
             against the measured HBM copy bandwidth (MEASURED_PEAKS.json).
  cpu_baseline  the reference's own software path (oracle/_ref: src/qatzip_sw.c + zlib) on the
             host cores of this box, a bounded sample of the same bytes, same parameters.

`--impl reference` times only that CPU path (all host threads) and prints the same JSON shape.
Inputs are larger than L2 (4 GiB >> 126 MB), so no explicit flush is needed between steps.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from harness import qzapi as q  # noqa: E402

GB = 1e9
CALL_BYTES = 512 << 20
CHUNK = 65536


def env_int(name, d):
    v = os.environ.get(name)
    return int(v) if v else d


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 100 ms while the timed regions run."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def kernel_source_sha():
    """identifies the deflate kernel sources an ncu capture belongs to"""
    import hashlib
    h = hashlib.sha256()
    for f in ("qz_deflate.cu", "qz_match.cuh", "qz_warp.cuh", "qz_huffman.h"):
        h.update(open(os.path.join(ROOT, "qatzip_b200", "csrc", f), "rb").read())
    return h.hexdigest()


def cpu_reference_pass(lib_path, host_addr, nbytes, threads):
    """The reference's software path over host_addr[0:nbytes): one session per thread, contiguous
    slices, simultaneous start (pattern of reference test/main.c:2175-2202).  Returns (seconds, out_bytes)."""
    ref = q.QzLib(lib_path)
    per = (nbytes // threads) // CHUNK * CHUNK
    per = max(per, CHUNK)
    threads = max(1, min(threads, nbytes // per))
    outs, barrier = [0] * threads, threading.Barrier(threads + 1)

    def work(t):
        sess = ref.new_session(fmt=q.QZ_DEFLATE_GZIP_EXT, level=1, hw_buff_sz=CHUNK)
        lo, hi = t * per, (t + 1) * per if t + 1 < threads else nbytes
        cap = ref.lib.qzMaxCompressedLength(min(hi - lo, CALL_BYTES), None)
        dst = (C.c_ubyte * cap)()
        barrier.wait()
        made_total = 0
        for off in range(lo, hi, CALL_BYTES):
            n = min(CALL_BYTES, hi - off)
            rc, used, made = ref.compress_call(sess, host_addr + off, n, C.addressof(dst), cap)
            assert rc == q.QZ_OK and used == n, (rc, used, n)
            made_total += made
        outs[t] = made_total
        barrier.wait()
        ref.end_session(sess)

    ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    [t.start() for t in ths]
    barrier.wait(); t0 = time.perf_counter()
    barrier.wait(); dt = time.perf_counter() - t0
    [t.join() for t in ths]
    return dt, sum(outs), threads


_REAL_STDOUT = None


def _claim_stdout():
    """stdout carries exactly one JSON line: whatever libraries print there meanwhile (NCCL's version banner at the first
    collective, for one) goes to stderr instead"""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(obj):
    data = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-threads", type=int, default=env_int("QZ_BENCH_E2E_THREADS", 1), help="submitting threads (one session each) of the e2e pass")
    ap.add_argument("--gib", type=float, default=float(os.environ.get("QZ_BENCH_GIB", "4")), help="per-GPU workload (GiB); 4 = BASELINE config")
    args = ap.parse_args()
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    nbytes = int(args.gib * (1 << 30)) // CALL_BYTES * CALL_BYTES or CALL_BYTES
    ncores = os.cpu_count() or 1
    ref_lib = q.REF_SO if os.path.exists(q.REF_SO) else None
    workload = f"qzCompress QZ_DEFLATE_GZIP_EXT L1 hw_buff_sz=64KiB, {nbytes / (1 << 30):g} GiB SILESIA-LIKE per GPU in 512 MiB calls"
    # the same dictionary in both arms (the reference arm times a bounded sample of this workload per step: cpu_baseline.sample)
    config = {"workload": workload, "format": "QZ_DEFLATE_GZIP_EXT", "level": 1, "hw_buff_sz": CHUNK, "per_gpu_bytes": nbytes,
              "l2": "inputs larger than L2 (4 GiB per GPU vs 126 MB)"}

    import __graft_entry__ as ge
    if not (os.path.exists(q.CORPUS_SO) and os.path.exists(q.PORT_SO)):
        ge.build_checkers()
    cor = q.Corpus()

    # ---------------------------------------------------------------- reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        sample = min(nbytes, max(1 << 30, ncores * (32 << 20)))
        buf = (C.c_ubyte * sample)()
        cor.fill(q.Corpus.SILESIA_LIKE, C.addressof(buf), sample, threads=min(64, ncores))
        lib = ref_lib or None
        assert lib, "oracle/_ref missing (build it where /root/reference exists)"
        times = []
        for i in range(args.warmup + args.steps):
            dt, out, thr = cpu_reference_pass(lib, C.addressof(buf), sample, ncores)
            if i >= args.warmup:
                times.append(dt)
        dt = sum(times) / len(times)
        v = sample / dt / GB
        _emit(({"impl": "reference", "metric": "qzCompress GB/s (input) at 64 KiB chunks", "value": round(v, 4), "unit": "GB/s",
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 2),
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                          "config": config,
                          "cpu_baseline": {"value": round(v, 4), "unit": "GB/s", "cores": thr, "kind": "reference",
                                           "sample": f"{sample >> 20} MiB of the workload per step, one slice per thread"},
                          "e2e": {"value": round(v, 4), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "ratio": round(out / sample, 4)}))
        return 0

    # ---------------------------------------------------------------- our arm (GPU)
    os.environ.setdefault("QZB200_DEVICE", str(local))
    prod = q.QzLib(q.PRODUCT_SO)
    L = prod.lib
    assert L.qzb200DeviceCount() > 0, "no CUDA device: libqatzip.so has no CPU path"
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist:
            dist.barrier()

    def allmax(x):
        if not dist:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t.item())

    def allsum(x):
        if not dist:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.SUM); return float(t.item())

    sess = prod.new_session(fmt=q.QZ_DEFLATE_GZIP_EXT, level=1, hw_buff_sz=CHUNK)
    # host input: pinned pages from qzMalloc, filled with this rank's shard of the corpus
    h_in = L.qzMalloc(nbytes, 0, q.PINNED_MEM)
    assert h_in, "qzMalloc(PINNED) failed"
    segs_per_rank = nbytes >> 20
    cor.fill(q.Corpus.SILESIA_LIKE, h_in, nbytes, first_seg=rank * segs_per_rank, threads=max(1, min(64, ncores // max(1, world))))
    out_cap_call = L.qzMaxCompressedLength(CALL_BYTES, None)
    h_out = L.qzMalloc(out_cap_call, 0, q.PINNED_MEM)
    d_in = L.qzb200DeviceAlloc(nbytes)
    d_out = L.qzb200DeviceAlloc(out_cap_call)
    assert h_out and d_in and d_out
    assert L.qzb200CopyToDevice(d_in, h_in, nbytes) == 0
    ncalls = nbytes // CALL_BYTES

    def device_pass():
        made_total, codec_ms, codec_launches, launches, kernel_ms = 0, 0.0, 0, 0, 0.0
        for i in range(ncalls):
            rc, used, made, _ = prod.compress_device(sess, d_in + i * CALL_BYTES, CALL_BYTES, d_out, out_cap_call, 1)
            assert rc == q.QZ_OK and used == CALL_BYTES, (rc, used)
            st = prod.stats(sess)
            made_total += made; codec_ms += st.codec_ms; codec_launches += st.codec_launches; launches += st.kernel_launches; kernel_ms += st.kernel_ms
        return made_total, codec_ms, codec_launches, launches, kernel_ms

    host_break = {"h2d_ms": 0.0, "kernel_ms": 0.0, "d2h_ms": 0.0}
    T = max(1, min(args.e2e_threads, ncalls))
    e2e_sess = [sess] + [prod.new_session(fmt=q.QZ_DEFLATE_GZIP_EXT, level=1, hw_buff_sz=CHUNK) for _ in range(T - 1)]
    e2e_out = [h_out] + [L.qzMalloc(out_cap_call, 0, q.PINNED_MEM) for _ in range(T - 1)]
    assert all(e2e_out)

    def host_pass(threads):
        """every 512 MiB call of the shard through qzCompress(host -> host); thread t issues calls t, t + threads, ..."""
        made = [0] * threads
        for k in host_break:
            host_break[k] = 0.0

        def work(t):
            for i in range(t, ncalls, threads):
                rc, used, m = prod.compress_call(e2e_sess[t], h_in + i * CALL_BYTES, CALL_BYTES, e2e_out[t], out_cap_call, 1)
                assert rc == q.QZ_OK and used == CALL_BYTES, (rc, used)
                made[t] += m
                if threads == 1:
                    st = prod.stats(e2e_sess[t])
                    host_break["h2d_ms"] += st.h2d_ms; host_break["kernel_ms"] += st.kernel_ms; host_break["d2h_ms"] += st.d2h_ms
        if threads == 1:
            work(0)
        else:
            ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
            [t.start() for t in ths]
            [t.join() for t in ths]
        return sum(made)

    for _ in range(args.warmup):
        device_pass()
    clocks = ClockSampler(local)
    barrier(); clocks.start(); t0 = time.perf_counter()
    made = codec_ms = kernel_ms = 0.0; codec_launches = launches = 0
    for _ in range(args.steps):
        m, cm, cl, ln, km = device_pass()
        made, codec_ms, codec_launches, launches, kernel_ms = m, codec_ms + cm, codec_launches + cl, launches + ln, kernel_ms + km
    barrier(); dt = allmax(time.perf_counter() - t0)
    ms_per_step = dt / args.steps * 1e3
    total_in = allsum(float(nbytes))
    value = total_in / (dt / args.steps) / GB

    # end to end through qzCompress with host buffers
    def timed_host(threads):
        for _ in range(min(args.warmup, 2)):
            host_pass(threads)
        barrier(); t0 = time.perf_counter()
        for _ in range(args.steps):
            made_h = host_pass(threads)
        barrier(); d = allmax(time.perf_counter() - t0)
        return d, made_h
    dt_1, made_h = timed_host(1)
    break_1 = dict(host_break)
    dt_e, made_h = timed_host(T) if T > 1 else (dt_1, made_h)
    clk = clocks.stop()
    e2e = total_in / (dt_e / args.steps) / GB
    e2e_1 = total_in / (dt_1 / args.steps) / GB

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        per_launch_bytes = (nbytes + made) / (codec_launches / args.steps)
        per_launch_s = codec_ms / 1e3 / codec_launches
        achieved = per_launch_bytes / per_launch_s / GB
        # DRAM bytes per launch come from an ncu capture (profiles/traffic.json, written by tools/collect_profiles.py); they are
        # reported only while the kernel sources are the ones that were captured
        traffic, traffic_note = None, "no capture"
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if tj.get("source_sha") == kernel_source_sha():
                traffic, traffic_note = tj.get("deflate_dram_bytes_per_launch"), f"ncu capture {tj.get('capture')} of these kernel sources ({tj.get('source_sha', '')[:12]})"
            else:
                traffic_note = f"capture {tj.get('capture')} is of other kernel sources ({str(tj.get('source_sha'))[:12]}): not reported"
        # CPU baseline: bounded sample of the same bytes through the reference's software path
        cpu = None
        if ref_lib and world == 1 and not os.environ.get("QZ_BENCH_NOCPU"):      # reported at N = 1 only
            sample = min(nbytes, max(256 << 20, ncores * (16 << 20)))
            dtc, outc, thr = cpu_reference_pass(ref_lib, h_in, sample, ncores)
            cpu = {"value": round(sample / dtc / GB, 4), "unit": "GB/s", "cores": thr, "kind": "reference",
                   "sample": f"first {sample >> 20} MiB of rank 0's shard, one pass, {thr} threads", "ratio": round(outc / sample, 4)}
        st = prod.stats(sess)
        line = {
            "metric": "qzCompress GB/s (input) at 64 KiB chunks", "value": round(value, 3), "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3),
            "kernel_ms_per_step_cuda_events": round(kernel_ms / args.steps, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": config,
            "kernel": {"deflate_blocks": "one per 64 KiB window (window kernel)" if st.group_blocks else "one per 8 KiB piece", "piece_log2": st.piece_log2},
            "ratio": round(made / nbytes, 4),
            "e2e": {"value": round(e2e, 3), "unit": "GB/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": int(made_h),
                    "api": f"qzCompress(host pinned -> host pinned), 512 MiB per call, {T} submitting thread(s) with one session each",
                    "threads": T, "ms_per_step": round(dt_e / args.steps * 1e3, 3),
                    "one_thread": {"value": round(e2e_1, 3), "ms_per_step": round(dt_1 / args.steps * 1e3, 3),
                                   "stage_ms_per_step_summed_overlapping": {k: round(v, 2) for k, v in break_1.items()}}},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "qzb_deflate_window_kernel" if st.group_blocks else "qzb_deflate_pieces_kernel", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_source": traffic_note, "peak_source": peak_src,
                         "bytes_per_launch": int(per_launch_bytes), "ms_per_launch": round(per_launch_s * 1e3, 4)},
            "cpu_baseline": cpu, "clocks": clk}
    # ---------------------------------------------------------------- secondary legs (BASELINE configs[2..4]), after the timed headline
    secondary = {}
    if os.environ.get("QZ_BENCH_SECONDARY", "1") != "0":
        from harness import bench_secondary as bs
        for p in (h_in, h_out, *e2e_out[1:]):
            L.qzFree(p)
        L.qzb200DeviceFree(d_in); L.qzb200DeviceFree(d_out)
        peak2 = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
        sec_steps = max(1, min(args.steps, 3))
        try:
            secondary["lz4_config3"] = bs.lz4_leg(prod, ref_lib, cor, peak2, ncores, sec_steps, rank, world, barrier, allmax, allsum)
        except Exception as e:          # a secondary leg never takes the headline down
            secondary["lz4_config3"] = {"error": repr(e)}
        # one process driving every visible GPU (the other ranks of a torchrun launch sit at the barrier meanwhile)
        barrier()
        if rank == 0 and L.qzb200DeviceCount() > 1:
            try:
                secondary["one_process_all_gpus"] = bs.one_process_leg(prod, cor, ncores, sec_steps, nbytes)
                secondary["one_process_all_gpus_lz4"] = bs.one_process_leg(prod, cor, ncores, sec_steps, min(nbytes, 2 << 30), fmt=q.FMT_LZ4)
            except Exception as e:
                secondary["one_process_all_gpus"] = {"error": repr(e)}
        barrier()
        if rank == 0 and world == 1:
            for name, fn in (("inflate_config2", lambda: bs.inflate_leg(prod, ref_lib, cor, peak2, ncores, sec_steps) if ref_lib else {"unavailable": "oracle/_ref missing"}),
                             ("stream_config4", lambda: bs.stream_leg(prod, ref_lib, cor, peak2, ncores))):
                try:
                    secondary[name] = fn()
                except Exception as e:
                    secondary[name] = {"error": repr(e)}
    if rank == 0:
        line["secondary"] = secondary
        _emit(line)
    if dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
