"""Where a piecemeal stream's time goes: the C driver over 64 MiB in 4 KiB submissions with the engine's per-batch timeline on."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["QZB200_TIMELINE"] = "1"
from harness import qzapi as q
prod = q.QzLib(q.PRODUCT_SO); L = prod.lib
SN = 64 << 20
h_in = L.qzMalloc(SN, 0, q.PINNED_MEM); q.Corpus().fill(q.Corpus.SILESIA_LIKE, h_in, SN, threads=16)
drv = C.CDLL(os.path.join(os.path.dirname(q.CORPUS_SO), "libqzdrive.so"))
drv.qzdrive_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint, C.c_void_p, C.c_uint, C.c_void_p, C.c_size_t,
                               C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint), C.POINTER(C.c_double)]
fn_s, fn_e = C.cast(L.qzCompressStream, C.c_void_p), C.cast(L.qzEndStream, C.c_void_p)
for rep in range(2):
    sess = prod.new_session(fmt=q.QZ_DEFLATE_RAW, strm_buff_sz=65536)
    ocap = 8 << 20; obuf = (C.c_ubyte * ocap)()
    outb, calls, crc, secs = C.c_uint64(0), C.c_uint64(0), C.c_uint(0), C.c_double(0)
    t0 = time.perf_counter()
    rc = drv.qzdrive_stream(fn_s, fn_e, C.byref(sess), h_in, SN, 4096, obuf, ocap, None, 0, C.byref(outb), C.byref(calls), C.byref(crc), C.byref(secs))
    sys.stderr.write(f"--- rep {rep}: rc {rc} {SN / secs.value / 1e6:.0f} MB/s loop {secs.value * 1e3:.1f} ms wall {1e3 * (time.perf_counter() - t0):.1f} ms\n")
    prod.end_session(sess)
