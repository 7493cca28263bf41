"""Secondary legs of bench.py: BASELINE.json configs[2], [3] and [4], measured after the headline in the same run.

  inflate  configs[2]  qzDecompress QZ_DEFLATE_GZIP over gzip members of 4..256 KiB (uncompressed) made by the REFERENCE's
                       software path (zlib level 1) from the SILESIA-LIKE corpus, QZ_BENCH_INFLATE_GIB (default 8) GiB of
                       output.  Rank 0 only.
  lz4      configs[3]  qzCompress QZ_LZ4, 64 KiB blocks, QZ_BENCH_LZ4_GIB (default 2) GiB per rank: at N = 8 the 16 GiB config.
                       Every rank; weak scaling like the headline.
  stream   configs[4]  qzCompressStream QZ_DEFLATE_RAW fed 4 KiB per call from C (harness/stream_drive.c),
                       QZ_BENCH_STREAM_GIB (default 1) GiB.  Rank 0 only.

Every leg reports `value` (device-resident, or for the stream leg the API itself), `e2e` (host pinned -> host pinned through
the C ABI), `roofline` of its dominant kernel (algorithmic bytes in + out per launch / CUDA-event time, against the measured
HBM copy bandwidth) and the reference's software path on the host cores over a bounded sample.  Harness code: it calls the
product only through libqatzip.so and the reference only through oracle/_ref.
"""
import ctypes as C
import os
import threading
import time

from harness import qzapi as q

GB = 1e9
CALL = 512 << 20


def _env_f(name, d):
    v = os.environ.get(name)
    return float(v) if v else d


def _roof(bytes_algo, kernel_ms, launches, peak):
    if not kernel_ms or not launches:
        return None
    a = bytes_algo / (kernel_ms / 1e3) / GB
    return {"bound": "hbm", "achieved": round(a, 2), "peak": peak, "unit": "GB/s", "frac": round(a / peak, 4),
            "bytes_per_launch": int(bytes_algo / launches), "ms_per_launch": round(kernel_ms / launches, 4), "launches": int(launches)}


def _threads_run(nthreads, fn):
    """run fn(t) on nthreads threads with a common start; returns seconds between the two barriers"""
    bar = threading.Barrier(nthreads + 1)

    def work(t):
        bar.wait(); fn(t); bar.wait()
    ths = [threading.Thread(target=work, args=(t,)) for t in range(nthreads)]
    [t.start() for t in ths]
    bar.wait(); t0 = time.perf_counter(); bar.wait(); dt = time.perf_counter() - t0
    [t.join() for t in ths]
    return dt


# ------------------------------------------------------------------------------------------------ configs[2]: inflate
def make_reference_members(ref_lib, h_src, n_out, ncores):
    """gzip members of 4..256 KiB made by the reference's software path over h_src[0:n_out) -> (bytes, [(off, len)], seconds)"""
    ref = q.QzLib(ref_lib)
    sizes, state, pos, cuts = [4, 8, 16, 32, 64, 128, 256], 3, 0, []
    while pos < n_out:
        state = (state * 6364136223846793005 + 1442695040888963407) & ((1 << 64) - 1)
        n = min(sizes[(state >> 33) % 7] << 10, n_out - pos)
        cuts.append((pos, n)); pos += n
    T = max(1, min(64, ncores))
    outs = [None] * len(cuts)

    def work(t):
        sess = ref.new_session(fmt=q.QZ_DEFLATE_GZIP, level=1, hw_buff_sz=262144)
        dst = (C.c_ubyte * (300 << 10))()
        for i in range(t, len(cuts), T):
            o, n = cuts[i]
            rc, used, made = ref.compress_call(sess, h_src + o, n, C.addressof(dst), len(dst))
            assert rc == 0 and used == n
            outs[i] = C.string_at(C.addressof(dst), made)
        ref.end_session(sess)
    dt = _threads_run(T, work)
    return outs, cuts, dt


def inflate_leg(prod, ref_lib, cor, peak, ncores, steps):
    L = prod.lib
    n_out = int(_env_f("QZ_BENCH_INFLATE_GIB", 8.0) * (1 << 30)) // (1 << 20) * (1 << 20)
    h_plain = L.qzMalloc(n_out, 0, q.PINNED_MEM)
    assert h_plain, "qzMalloc for the inflate leg failed"
    cor.fill(q.Corpus.SILESIA_LIKE, h_plain, n_out, threads=max(1, min(64, ncores)))
    members, cuts, make_s = make_reference_members(ref_lib, h_plain, n_out, ncores)
    total_c = sum(len(m) for m in members)
    h_c = L.qzMalloc(total_c, 0, q.PINNED_MEM)
    h_back = L.qzMalloc(n_out, 0, q.PINNED_MEM)
    assert h_c and h_back
    # calls of at most 3.5 GiB of output (the API's lengths are 32-bit): [(c_off, c_len, out_off, out_len)]
    OUT_CALL = 3584 << 20
    calls, c_off, o_off, cl, ol = [], 0, 0, 0, 0
    off = 0
    for m, (_, n) in zip(members, cuts):
        C.memmove(h_c + off, m, len(m)); off += len(m)
        if ol + n > OUT_CALL:
            calls.append((c_off, cl, o_off, ol)); c_off += cl; o_off += ol; cl = ol = 0
        cl += len(m); ol += n
    calls.append((c_off, cl, o_off, ol))
    d_c, d_out = L.qzb200DeviceAlloc(max(c[1] for c in calls)), L.qzb200DeviceAlloc(max(c[3] for c in calls))
    assert d_c and d_out
    sess = prod.new_session(fmt=q.QZ_DEFLATE_GZIP, hw_buff_sz=262144)

    def device_pass():
        kms = launches = 0
        wall = 0.0
        for (co, cn, oo, on) in calls:
            assert L.qzb200CopyToDevice(d_c, h_c + co, cn) == 0          # staging outside the timed part
            used, made = C.c_uint64(0), C.c_uint64(0)
            t0 = time.perf_counter()
            rc = L.qzb200DecompressDevice(C.byref(sess), d_c, h_c + co, cn, d_out, on, C.byref(used), C.byref(made))
            wall += time.perf_counter() - t0
            assert rc == 0 and used.value == cn and made.value == on, (rc, used.value, cn, made.value, on)
            st = prod.stats(sess)
            kms += st.kernel_ms; launches += st.kernel_launches
        return wall, kms, launches

    def host_pass():
        t0 = time.perf_counter()
        for (co, cn, oo, on) in calls:
            rc, used, made = prod.decompress_call(sess, h_c + co, cn, h_back + oo, on)
            assert rc == 0 and used == cn and made == on, (rc, used, cn, made, on)
        return time.perf_counter() - t0

    device_pass()
    best = min((device_pass() for _ in range(max(1, steps))), key=lambda r: r[0])
    host_pass()
    e2e_s = min(host_pass() for _ in range(max(1, steps)))
    exact = all(C.string_at(h_back + o, min(64 << 20, n_out - o)) == C.string_at(h_plain + o, min(64 << 20, n_out - o)) for o in range(0, n_out, 64 << 20))
    assert exact, "inflate leg: output differs from the input the reference compressed"
    prod.end_session(sess)
    # the reference's own inflate over a bounded sample of the same members, all host threads
    ref = q.QzLib(ref_lib)
    T = max(1, min(64, ncores))
    sample_members = min(len(members), max(T, int(len(members) * min(1.0, (1 << 30) / n_out))))
    per = (sample_members + T - 1) // T
    offs = [0]
    for m in members[:sample_members]:
        offs.append(offs[-1] + len(m))
    sample_out = sum(n for _, n in cuts[:sample_members])

    def ref_work(t):
        lo, hi = t * per, min(sample_members, (t + 1) * per)
        if lo >= hi:
            return
        sess_r = ref.new_session(fmt=q.QZ_DEFLATE_GZIP, hw_buff_sz=262144)
        n = sum(c[1] for c in cuts[lo:hi])
        dst = (C.c_ubyte * n)()
        rc, used, made = ref.decompress_call(sess_r, h_c + offs[lo], offs[hi] - offs[lo], C.addressof(dst), n)
        assert rc == 0 and made == n, (rc, made, n)
        ref.end_session(sess_r)
    ref_s = _threads_run(T, ref_work)
    for p in (h_plain, h_c, h_back):
        L.qzFree(p)
    L.qzb200DeviceFree(d_c); L.qzb200DeviceFree(d_out)
    wall, kms, launches = best
    return {"config": "qzDecompress QZ_DEFLATE_GZIP, gzip members of 4-256 KiB made by the reference's software path (zlib level 1), "
                      f"{n_out / (1 << 30):g} GiB of output, SILESIA-LIKE",
            "metric": "qzDecompress GB/s (output)", "unit": "GB/s", "members": len(members), "compressed_ratio": round(total_c / n_out, 4),
            "value": round(n_out / wall / GB, 3), "value_kernels_only": round(n_out / (kms / 1e3) / GB, 3),
            "e2e": {"value": round(n_out / e2e_s / GB, 3), "unit": "GB/s", "h2d_bytes_per_step": total_c, "d2h_bytes_per_step": n_out,
                    "api": "qzDecompress(host pinned -> host pinned), calls of <= 3.5 GiB of output"},
            "bit_exact": exact, "roofline": dict(_roof(total_c + n_out, kms, launches, peak) or {}, kernel="qzb_inflate_kernel"),
            "cpu_baseline": {"value": round(sample_out / ref_s / GB, 3), "unit": "GB/s", "cores": T, "kind": "reference",
                             "sample": f"the first {sample_members} members ({sample_out >> 20} MiB of output), one slice per thread"},
            "members_made_by_reference_in_s": round(make_s, 2)}


# ------------------------------------------------------------------------------------------------ configs[3]: LZ4
def lz4_leg(prod, ref_lib, cor, peak, ncores, steps, rank, world, barrier, allmax, allsum):
    L = prod.lib
    nbytes = int(_env_f("QZ_BENCH_LZ4_GIB", 2.0) * (1 << 30)) // CALL * CALL or CALL
    h_in = L.qzMalloc(nbytes, 0, q.PINNED_MEM)
    cap = L.qzMaxCompressedLength(CALL, None)
    h_out = L.qzMalloc(cap, 0, q.PINNED_MEM)
    d_in, d_out = L.qzb200DeviceAlloc(nbytes), L.qzb200DeviceAlloc(cap)
    assert h_in and h_out and d_in and d_out
    cor.fill(q.Corpus.SILESIA_LIKE, h_in, nbytes, first_seg=rank * (nbytes >> 20), threads=max(1, min(64, ncores // max(1, world))))
    assert L.qzb200CopyToDevice(d_in, h_in, nbytes) == 0
    sess = prod.new_session(fmt=q.FMT_LZ4, hw_buff_sz=65536)
    ncalls = nbytes // CALL

    def device_pass():
        made = kms = cms = 0
        cl = 0
        for i in range(ncalls):
            rc, used, m, _ = prod.compress_device(sess, d_in + i * CALL, CALL, d_out, cap, 1)
            assert rc == 0 and used == CALL
            st = prod.stats(sess)
            made += m; kms += st.kernel_ms; cms += st.codec_ms; cl += st.codec_launches
        return made, kms, cms, cl

    def host_pass():
        made = 0
        for i in range(ncalls):
            rc, used, m = prod.compress_call(sess, h_in + i * CALL, CALL, h_out, cap, 1)
            assert rc == 0 and used == CALL
            made += m
        return made
    device_pass()
    barrier(); t0 = time.perf_counter()
    for _ in range(steps):
        made, kms, cms, cl = device_pass()
    barrier(); dt = allmax(time.perf_counter() - t0) / steps
    host_pass()
    barrier(); t0 = time.perf_counter()
    for _ in range(steps):
        made_h = host_pass()
    barrier(); dte = allmax(time.perf_counter() - t0) / steps
    total = allsum(float(nbytes))
    out = {"config": f"qzCompress QZ_LZ4, one 64 KiB block per 64 KiB chunk, {nbytes / (1 << 30):g} GiB SILESIA-LIKE per GPU ({total / (1 << 30):g} GiB in all), sharded by rank",
           "metric": "qzCompress GB/s (input)", "unit": "GB/s", "n_gpus": world, "value": round(total / dt / GB, 3), "ratio": round(made / nbytes, 4),
           "e2e": {"value": round(total / dte / GB, 3), "unit": "GB/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": int(made_h),
                   "api": "qzCompress(host pinned -> host pinned), 512 MiB per call"},
           "roofline": dict(_roof(nbytes + made, cms, cl, peak) or {}, kernel="qzb_lz4_window_kernel")}
    if rank == 0 and ref_lib and world == 1:
        ref = q.QzLib(ref_lib)
        T = max(1, min(64, ncores))
        sample = min(nbytes, max(256 << 20, T * (16 << 20)))
        per = sample // T // 65536 * 65536
        outs = [0] * T

        def work(t):
            s = ref.new_session(fmt=q.FMT_LZ4, hw_buff_sz=65536)
            dst = (C.c_ubyte * (70 << 10))()
            for o in range(t * per, (t + 1) * per, 65536):                 # one frame per 64 KiB chunk, like the hardware path
                rc, used, m = ref.compress_call(s, h_in + o, 65536, C.addressof(dst), len(dst))
                assert rc == 0
                outs[t] += m
            ref.end_session(s)
        ref_s = _threads_run(T, work)
        out["cpu_baseline"] = {"value": round(per * T / ref_s / GB, 3), "unit": "GB/s", "cores": T, "kind": "reference", "ratio": round(sum(outs) / (per * T), 4),
                               "sample": f"the first {per * T >> 20} MiB of rank 0's shard, one LZ4 frame per 64 KiB chunk"}
    prod.end_session(sess)
    L.qzFree(h_in); L.qzFree(h_out); L.qzb200DeviceFree(d_in); L.qzb200DeviceFree(d_out)
    return out


# ------------------------------------------------------------------------------------------------ configs[4]: stream
def stream_leg(prod, ref_lib, cor, peak, ncores):
    L = prod.lib
    n = int(_env_f("QZ_BENCH_STREAM_GIB", 1.0) * (1 << 30))
    drv_path = os.path.join(os.path.dirname(q.CORPUS_SO), "libqzdrive.so")
    if not os.path.exists(drv_path):
        return {"unavailable": "harness/libqzdrive.so not built"}
    drv = C.CDLL(drv_path)
    drv.qzdrive_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint, C.c_void_p, C.c_uint, C.c_void_p, C.c_size_t,
                                   C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint), C.POINTER(C.c_double)]
    h_in = L.qzMalloc(n, 0, q.PINNED_MEM)
    assert h_in
    cor.fill(q.Corpus.SILESIA_LIKE, h_in, n, threads=max(1, min(64, ncores)))

    def run(lib, nbytes):
        sess = lib.new_session(fmt=q.QZ_DEFLATE_RAW, strm_buff_sz=65536)
        ocap = 16 << 20
        obuf = (C.c_ubyte * ocap)()
        outb, calls, crc, secs = C.c_uint64(0), C.c_uint64(0), C.c_uint(0), C.c_double(0)
        rc = drv.qzdrive_stream(C.cast(lib.lib.qzCompressStream, C.c_void_p), C.cast(lib.lib.qzEndStream, C.c_void_p), C.byref(sess), h_in, nbytes, 4096,
                                obuf, ocap, None, 0, C.byref(outb), C.byref(calls), C.byref(crc), C.byref(secs))
        assert rc == 0, rc
        kms = launches = None
        if lib is prod:
            st = prod.stats(sess); kms, launches = st.kernel_ms, st.kernel_launches
        lib.end_session(sess)
        return secs.value, outb.value, calls.value, crc.value
    run(prod, min(n, 64 << 20))
    # host-bound and noisy (one thread copying 4 KiB pieces): the better of two passes, both reported
    passes = [run(prod, n) for _ in range(2)]
    secs, outb, calls, crc = min(passes, key=lambda r: r[0])
    import zlib
    crc_ok = crc == (zlib.crc32(C.string_at(h_in, min(n, 1 << 30))) & 0xffffffff) if n <= (1 << 30) else None
    out = {"config": f"qzCompressStream QZ_DEFLATE_RAW, 4 KiB submissions from C, {n / (1 << 30):g} GiB stream, strm_buff_sz 64 KiB",
           "metric": "qzCompressStream GB/s (input)", "unit": "GB/s", "value": round(n / secs / GB, 3), "calls": int(calls), "ratio": round(outb / n, 4),
           "stream_crc_matches_zlib": crc_ok, "passes_GBps": [round(n / r[0] / GB, 3) for r in passes],
           "e2e": {"value": round(n / secs / GB, 3), "unit": "GB/s", "h2d_bytes_per_step": n, "d2h_bytes_per_step": int(outb),
                   "api": "qzCompressStream(host memory in, host memory out): the value IS the end-to-end number"},
           "roofline": None}
    if hasattr(drv, "qzdrive_copy_ceiling"):
        # the same loop with no compressor behind it: 4 KiB copies into a pinned staging buffer, the compressed share copied out
        drv.qzdrive_copy_ceiling.argtypes = [C.c_void_p, C.c_size_t, C.c_uint, C.c_void_p, C.c_void_p, C.c_uint, C.c_uint, C.c_void_p, C.POINTER(C.c_double)]
        cap = 8 << 20
        s_in, s_out = L.qzMalloc(cap, 0, q.PINNED_MEM), L.qzMalloc(cap, 0, q.PINNED_MEM)
        obuf = (C.c_ubyte * cap)()
        cs = C.c_double(0)
        best = None
        for _ in range(3):
            drv.qzdrive_copy_ceiling(h_in, n, 4096, s_in, s_out, cap, int(cap * outb / n), obuf, C.byref(cs))
            best = cs.value if best is None else min(best, cs.value)
        L.qzFree(s_in); L.qzFree(s_out)
        out["host_copy_ceiling"] = {"value": round(n / best / GB, 3), "unit": "GB/s",
                                    "what": "the submission loop with memcpy only (4 KiB in to pinned staging, compressed share out), one thread: bounds any stream API here"}
        out["frac_of_host_copy_ceiling"] = round(out["value"] / out["host_copy_ceiling"]["value"], 3)
    if ref_lib:
        ref = q.QzLib(ref_lib)
        sample = min(n, 64 << 20)
        rs, ro, rc_, _ = run(ref, sample)
        out["cpu_baseline"] = {"value": round(sample / rs / GB, 3), "unit": "GB/s", "cores": 1, "kind": "reference", "ratio": round(ro / sample, 4),
                               "sample": f"the first {sample >> 20} MiB of the stream through the reference's qzCompressStream (one thread: a stream is serial)"}
    L.qzFree(h_in)
    return out


# ------------------------------------------------------------------------------------------------ one process, all GPUs
def one_process_leg(prod, cor, ncores, steps, nbytes, fmt=q.QZ_DEFLATE_GZIP_EXT):
    """the headline e2e call pattern (qzCompress, host pinned -> host pinned, 512 MiB calls) from ONE process with
    QZB200_DEVICES=all: (a) one submitting thread -- every call deals its batches over all visible GPUs (per-GPU submission
    queues inside the library); (b) one submitting thread per GPU with a session each (the reference's own benchmark pattern,
    test/main.c:2175-2202) -- the sessions take the GPUs in turn as their primary device and still spread every call"""
    L = prod.lib
    ndev = L.qzb200DeviceCount()
    ncalls = max(ndev, nbytes // CALL)
    nbytes = ncalls * CALL
    h_in = L.qzMalloc(nbytes, 0, q.PINNED_MEM)
    cap = L.qzMaxCompressedLength(CALL, None)
    assert h_in
    cor.fill(q.Corpus.SILESIA_LIKE, h_in, nbytes, threads=max(1, min(64, ncores)))
    old = os.environ.get("QZB200_DEVICES")
    os.environ["QZB200_DEVICES"] = "all"
    res = {}
    for label, T in (("one_thread", 1), ("thread_per_gpu", ndev)):
        sess = [prod.new_session(fmt=fmt, hw_buff_sz=65536) for _ in range(T)]
        outs = [L.qzMalloc(cap, 0, q.PINNED_MEM) for _ in range(T)]
        assert all(outs)
        made = [0] * T

        def work(t):
            made[t] = 0
            for i in range(t, ncalls, T):
                rc, used, m = prod.compress_call(sess[t], h_in + i * CALL, CALL, outs[t], cap, 1)
                assert rc == 0 and used == CALL
                made[t] += m
        _threads_run(T, work)
        dt = min(_threads_run(T, work) for _ in range(steps))
        devices = prod.stats(sess[0]).devices
        for s_ in sess:
            prod.end_session(s_)
        for o in outs:
            L.qzFree(o)
        res[label] = {"value": round(nbytes / dt / GB, 3), "unit": "GB/s", "threads": T, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": int(sum(made)),
                      "api": "qzCompress(host pinned -> host pinned), 512 MiB per call, one session per thread"}
    if old is None:
        os.environ.pop("QZB200_DEVICES", None)
    else:
        os.environ["QZB200_DEVICES"] = old
    L.qzFree(h_in)
    return {"config": f"qzCompress {'QZ_LZ4' if fmt == q.FMT_LZ4 else 'QZ_DEFLATE_GZIP_EXT'}, 64 KiB chunks, {nbytes / (1 << 30):g} GiB, ONE process, QZB200_DEVICES=all",
            "devices": int(devices), "visible_devices": int(ndev), "metric": "qzCompress GB/s (input), end to end", "unit": "GB/s",
            "e2e": res["thread_per_gpu"], "e2e_one_thread": res["one_thread"], "ratio": round(sum(made) / (nbytes / 1) if False else res["thread_per_gpu"]["d2h_bytes_per_step"] / nbytes, 4)}
