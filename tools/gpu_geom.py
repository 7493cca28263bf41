"""Device-resident deflate timing for one launch geometry (QZB200_WARPS / QZB200_BUFFERS from the environment)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from harness import qzapi as q
prod = q.QzLib(q.PRODUCT_SO); L = prod.lib
n = 512 << 20
h = L.qzMalloc(n, 0, q.PINNED_MEM); q.Corpus().fill(q.Corpus.SILESIA_LIKE, h, n, threads=16)
cap = L.qzMaxCompressedLength(n, None)
d_in, d_out = L.qzb200DeviceAlloc(n), L.qzb200DeviceAlloc(cap)
assert L.qzb200CopyToDevice(d_in, h, n) == 0
sess = prod.new_session(fmt=q.QZ_DEFLATE_GZIP_EXT, level=int(os.environ.get("GEOM_LEVEL", "1")), hw_buff_sz=65536)
ms = []
for it in range(6):
    rc, used, made, _ = prod.compress_device(sess, d_in, n, d_out, cap, 1)
    assert rc == 0
    ms.append(prod.stats(sess).codec_ms)
print(json.dumps({"warps": os.environ.get("QZB200_WARPS"), "bufs": os.environ.get("QZB200_BUFFERS"), "codec_ms_min": round(min(ms[2:]), 3), "ratio": round(made / n, 4)}))
