"""Per-batch timeline of one synchronous 512 MiB qzCompress call with host pinned buffers (QZB200_TIMELINE=1 makes the
engine print device event times and host wake-up times for every batch on stderr): where a call's milliseconds go."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["QZB200_TIMELINE"] = "1"
from harness import qzapi as q
prod = q.QzLib(q.PRODUCT_SO); L = prod.lib
n = 512 << 20
h = L.qzMalloc(n, 0, q.PINNED_MEM); q.Corpus().fill(q.Corpus.SILESIA_LIKE, h, n, threads=16)
cap = L.qzMaxCompressedLength(n, None)
o = L.qzMalloc(cap, 0, q.PINNED_MEM)
sess = prod.new_session(fmt=q.QZ_DEFLATE_GZIP_EXT, level=1, hw_buff_sz=65536)
for it in range(3):
    sys.stderr.write(f"--- call {it}\n"); sys.stderr.flush()
    t0 = time.perf_counter()
    rc, used, made = prod.compress_call(sess, h, n, o, cap, 1)
    dt = time.perf_counter() - t0
    assert rc == 0 and used == n
    sys.stderr.write(f"--- call {it}: {dt * 1e3:.2f} ms wall, {n / dt / 1e9:.1f} GB/s, out {made}\n")
