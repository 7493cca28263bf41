"""Per-phase shares of warp time in the deflate kernel in use (window kernel by default, QZB200_WINDOW=0: per piece) (needs the A/B build:
make -C qatzip_b200/csrc ab ABFLAGS=-DQZ_PHASE_CLOCKS).  Prints lane-0 cycles per phase, summed over warps."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["QZ_PRODUCT_SO"] = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "qatzip_b200", "libqatzip_ab.so")
from harness import qzapi as q
prod = q.QzLib(q.PRODUCT_SO); L = prod.lib
n = 512 << 20
h = L.qzMalloc(n, 0, q.PINNED_MEM); q.Corpus().fill(q.Corpus.SILESIA_LIKE, h, n, threads=32)
cap = L.qzMaxCompressedLength(n, None)
d_in, d_out = L.qzb200DeviceAlloc(n), L.qzb200DeviceAlloc(cap)
assert L.qzb200CopyToDevice(d_in, h, n) == 0
sess = prod.new_session(fmt=q.QZ_DEFLATE_GZIP_EXT, level=1, hw_buff_sz=65536)
names = ["ticket + unit/buffer wait", "load+crc", "match+select", "slot histogram", "sort", "huffman lengths", "header plan+cost", "codes+prefix", "emit",
         "wait: slowest matcher / coder", "coder: waiting for the matchers", "leader: header emit + tables", "count pass", "wait: bit totals", "wait: zeroed words", "prepass + seed (with their barriers)"]
out = (C.c_ulonglong * 16)()
for it in range(3):
    L.qzb_phase_cycles_read(out, 1)
    rc, used, made, _ = prod.compress_device(sess, d_in, n, d_out, cap, 1)
    st = prod.stats(sess)
    L.qzb_phase_cycles_read(out, 0)
tot = sum(out[:16])
res = {names[i]: round(out[i] / tot, 4) for i in range(16) if out[i]}
res["codec_ms"] = st.codec_ms; res["ratio"] = made / n
print(json.dumps(res))
