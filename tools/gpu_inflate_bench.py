"""Inflate measurements on the GPU box: (a) 64 KiB gzip-ext members made by our compressor, (b) BASELINE
configs[2] in small: gzip members made by the reference software path (zlib -1), uncompressed sizes
drawn from {4..256} KiB.  Device-resident decode, CUDA-event kernel time from qzb200GetStats."""
import ctypes as C, json, os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from harness import qzapi as q
prod, cor = q.QzLib(q.PRODUCT_SO), q.Corpus()
L = prod.lib
N = int(os.environ.get("INFL_MIB", "512")) << 20
REPS = int(os.environ.get("INFL_REPS", "4"))
h_in = L.qzMalloc(N, 0, q.PINNED_MEM); cor.fill(q.Corpus.SILESIA_LIKE, h_in, N, threads=16)
cap = L.qzMaxCompressedLength(N, None)
h_c = L.qzMalloc(cap, 0, q.PINNED_MEM); h_back = L.qzMalloc(N, 0, q.PINNED_MEM)
d_in, d_c, d_back = L.qzb200DeviceAlloc(N), L.qzb200DeviceAlloc(cap), L.qzb200DeviceAlloc(N)
L.qzb200CopyToDevice(d_in, h_in, N)
res = {}

def decode(name, fmt, hw, total_c, n_out):
    sess = prod.new_session(fmt=fmt, hw_buff_sz=hw)
    best = 1e9
    for rep in range(REPS):
        used, made = C.c_uint64(0), C.c_uint64(0)
        t0 = time.perf_counter()
        rc = L.qzb200DecompressDevice(C.byref(sess), d_c, h_c, total_c, d_back, n_out, C.byref(used), C.byref(made))
        wall = time.perf_counter() - t0
        assert rc == 0 and used.value == total_c and made.value == n_out, (rc, used.value, total_c, made.value)
        st = prod.stats(sess)
        best = min(best, st.kernel_ms)
    L.qzb200CopyToHost(h_back, d_back, n_out)
    ok = all(C.string_at(h_back + o, min(256 << 20, n_out - o)) == C.string_at(h_in + o, min(256 << 20, n_out - o)) for o in range(0, n_out, 256 << 20))
    res[name] = {"GBps_out_kernels": round(n_out / (best / 1e3) / 1e9, 2), "kernel_ms": round(best, 3), "GBps_out_wall_last": round(n_out / wall / 1e9, 2),
                 "members": int(st.units), "launches": int(st.kernel_launches), "exact": ok, "MiB_out": n_out >> 20}
    prod.end_session(sess)
    assert ok

if "ours" in os.environ.get("INFL_CASES", "ours,ref"):
    sess = prod.new_session(fmt=q.QZ_DEFLATE_GZIP_EXT)
    rc, used, made, _ = prod.compress_device(sess, d_in, N, d_c, cap, 1); assert rc == 0
    prod.end_session(sess)
    L.qzb200CopyToHost(h_c, d_c, made)
    decode("ours_64KiB_gzip_ext", q.QZ_DEFLATE_GZIP_EXT, 65536, made, N)

if "ref" in os.environ.get("INFL_CASES", "ours,ref") and os.path.exists(q.REF_SO):
    ref = q.QzLib(q.REF_SO)
    M = min(N, int(os.environ.get("INFL_REF_MIB", "256")) << 20)
    sizes, state, pos, cuts = [4, 8, 16, 32, 64, 128, 256], 3, 0, []
    while pos + (256 << 10) <= M:
        state = (state * 6364136223846793005 + 1442695040888963407) & ((1 << 64) - 1)
        n = sizes[(state >> 33) % 7] << 10
        cuts.append((pos, n)); pos += n
    n_out = pos
    T = min(32, os.cpu_count() or 1)
    outs = [None] * len(cuts)
    def work(t):
        sess = ref.new_session(fmt=q.QZ_DEFLATE_GZIP, level=1, hw_buff_sz=262144)
        dst = (C.c_ubyte * (300 << 10))()
        for i in range(t, len(cuts), T):
            o, n = cuts[i]
            rc, used, made = ref.compress_call(sess, h_in + o, n, C.addressof(dst), len(dst))
            assert rc == 0 and used == n
            outs[i] = C.string_at(C.addressof(dst), made)
        ref.end_session(sess)
    ths = [threading.Thread(target=work, args=(t,)) for t in range(T)]
    [t.start() for t in ths]; [t.join() for t in ths]
    blob = b"".join(outs)
    C.memmove(h_c, blob, len(blob)); L.qzb200CopyToDevice(d_c, h_c, len(blob))
    res["ref_ratio"] = round(len(blob) / n_out, 4)
    decode("ref_mixed_4_256KiB_gzip", q.QZ_DEFLATE_GZIP, 262144, len(blob), n_out)
print(json.dumps(res))
