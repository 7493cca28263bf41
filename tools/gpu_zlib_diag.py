"""Timing + debug of the zlib-format paths on the GPU box (development aid)."""
import os, sys, time, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from harness import qzapi as q

prod, ref, cor = q.QzLib(q.PRODUCT_SO), q.QzLib(q.REF_SO), q.Corpus()
data = cor.make(q.Corpus.SILESIA_LIKE, 12 << 20)


def timed(label, fn):
    t0 = time.perf_counter()
    try:
        r = fn()
        print(f"{label}: {1e3 * (time.perf_counter() - t0):.1f} ms", flush=True)
        return r
    except AssertionError as e:
        print(f"{label}: FAILED after {1e3 * (time.perf_counter() - t0):.1f} ms: {e}", flush=True)
        return None


for n in (0, 1, 4096, 65536, 1060921, 3 << 20):
    d = data[:n]
    blob = timed(f"compress n={n}", lambda: prod.compress(d, fmt=q.FMT_ZLIB))
    if blob is not None and n:
        out = timed(f"  decompress ours n={n} ({len(blob)} B)", lambda: prod.decompress(blob, n + 8, fmt=q.FMT_ZLIB))
        assert out is None or out == d
    if n:
        rb = ref.compress(d, fmt=q.FMT_ZLIB)
        out = timed(f"  decompress reference-made n={n} ({len(rb)} B)", lambda: prod.decompress(rb, n + 8, fmt=q.FMT_ZLIB))
        assert out is None or out == d
for hw in (1024, 4096):
    d = data[:600000]
    blob = timed(f"compress hw={hw}", lambda: prod.compress(d, fmt=q.FMT_ZLIB, hw_buff_sz=hw))
    timed(f"  decompress hw={hw}", lambda: prod.decompress(blob, len(d) + 8, fmt=q.FMT_ZLIB, hw_buff_sz=hw))
os.environ["QZB200_BATCH_MB"] = "1"
os.environ["QZB200_DEBUG"] = "1"
big = data
bb = timed("compress 12 MiB (1 MiB batches)", lambda: prod.compress(big, fmt=q.FMT_ZLIB))
out = timed("  decompress in 4 MiB windows", lambda: prod.decompress(bb, len(big) + 8, fmt=q.FMT_ZLIB))
print("  equal:", out == big)
one = zlib.compress(big, 1)
out = timed("  one stream + 192 streams", lambda: prod.decompress(one + bb, 2 * len(big) + 8, fmt=q.FMT_ZLIB))
print("  equal:", out == big + big)
