"""Host-side stages of a device-resident qzb200DecompressDevice over mixed gzip members (QZB200_TIMELINE=1)."""
import ctypes as C, os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["QZB200_TIMELINE"] = "1"
from harness import qzapi as q
prod, cor, ref = q.QzLib(q.PRODUCT_SO), q.Corpus(), q.QzLib(q.REF_SO)
L = prod.lib
N = int(os.environ.get("TL_MIB", "2048")) << 20
h_in = L.qzMalloc(N, 0, q.PINNED_MEM); cor.fill(q.Corpus.SILESIA_LIKE, h_in, N, threads=16)
sizes, state, pos, cuts = [4, 8, 16, 32, 64, 128, 256], 3, 0, []
while pos + (256 << 10) <= N:
    state = (state * 6364136223846793005 + 1442695040888963407) & ((1 << 64) - 1)
    n = sizes[(state >> 33) % 7] << 10
    cuts.append((pos, n)); pos += n
n_out = pos
T = min(16, os.cpu_count() or 1); outs = [None] * len(cuts)
def work(t):
    sess = ref.new_session(fmt=q.QZ_DEFLATE_GZIP, level=1, hw_buff_sz=262144)
    dst = (C.c_ubyte * (300 << 10))()
    for i in range(t, len(cuts), T):
        o, n = cuts[i]
        rc, used, made = ref.compress_call(sess, h_in + o, n, C.addressof(dst), len(dst)); assert rc == 0
        outs[i] = C.string_at(C.addressof(dst), made)
    ref.end_session(sess)
ths = [threading.Thread(target=work, args=(t,)) for t in range(T)]; [t.start() for t in ths]; [t.join() for t in ths]
blob = b"".join(outs)
h_c = L.qzMalloc(len(blob) + 64, 0, q.PINNED_MEM); C.memmove(h_c, blob, len(blob))
d_c, d_back = L.qzb200DeviceAlloc(len(blob) + 64), L.qzb200DeviceAlloc(n_out)
L.qzb200CopyToDevice(d_c, h_c, len(blob))
sess = prod.new_session(fmt=q.QZ_DEFLATE_GZIP, hw_buff_sz=262144)
for rep in range(3):
    sys.stderr.write(f"--- rep {rep}\n")
    used, made = C.c_uint64(0), C.c_uint64(0)
    t0 = time.perf_counter()
    rc = L.qzb200DecompressDevice(C.byref(sess), d_c, h_c, len(blob), d_back, n_out, C.byref(used), C.byref(made))
    dt = time.perf_counter() - t0
    assert rc == 0 and made.value == n_out, (rc, used.value, made.value)
    sys.stderr.write(f"--- rep {rep}: {n_out / dt / 1e9:.2f} GB/s out, {dt * 1e3:.1f} ms wall, kernel {prod.stats(sess).kernel_ms:.1f} ms, {len(cuts)} members\n")
