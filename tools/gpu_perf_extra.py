"""Secondary measurements on the GPU box (not the headline): decompress, LZ4, stream API.
Prints one JSON object; device-resident numbers use CUDA-event kernel time from qzb200GetStats."""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from harness import qzapi as q

prod, cor = q.QzLib(q.PRODUCT_SO), q.Corpus()
L = prod.lib
N = int(os.environ.get("EXTRA_MIB", "1024")) << 20
CALL = 512 << 20
h_in = L.qzMalloc(N, 0, q.PINNED_MEM); cor.fill(q.Corpus.SILESIA_LIKE, h_in, N)
cap = L.qzMaxCompressedLength(CALL, None)
h_out = L.qzMalloc(cap * (N // CALL), 0, q.PINNED_MEM); h_back = L.qzMalloc(N, 0, q.PINNED_MEM)
d_in, d_out, d_back = L.qzb200DeviceAlloc(N), L.qzb200DeviceAlloc(cap * (N // CALL)), L.qzb200DeviceAlloc(N)
L.qzb200CopyToDevice(d_in, h_in, N)
res = {}
# raw PCIe copy bandwidth with the same pinned pages (what bounds the end-to-end number)
for nm, fn, a, b in (("h2d", L.qzb200CopyToDevice, d_in, h_in), ("d2h", L.qzb200CopyToHost, h_back, d_in)):
    fn(a, b, N); t0 = time.perf_counter(); fn(a, b, N); res["pcie_" + nm + "_GBps"] = round(N / (time.perf_counter() - t0) / 1e9, 2)
ONLY = os.environ.get("EXTRA_ONLY")
for name, fmt in (("gzip_ext", q.QZ_DEFLATE_GZIP_EXT), ("lz4", q.FMT_LZ4), ("zlib", q.FMT_ZLIB)):
    if ONLY and name != ONLY: continue
    sess = prod.new_session(fmt=fmt)
    # compress each 512 MiB call into consecutive regions; remember sizes
    sizes, kms = [], 0.0
    for rep in range(2):
        sizes, kms, off = [], 0.0, 0
        t0 = time.perf_counter()
        for i in range(N // CALL):
            rc, used, made, _ = prod.compress_device(sess, d_in + i * CALL, CALL, d_out + off, cap, 1)
            assert rc == 0 and used == CALL
            kms += prod.stats(sess).kernel_ms; sizes.append(made); off += made
        wall = time.perf_counter() - t0
    total_c = sum(sizes)
    res[name + "_compress"] = {"GBps_wall": round(N / wall / 1e9, 2), "GBps_kernels": round(N / (kms / 1e3) / 1e9, 2), "ratio": round(total_c / N, 4)}
    # host copy of the compressed stream for header walking, then device-resident decompress
    L.qzb200CopyToHost(h_out, d_out, total_c)
    for rep in range(2):
        kms, off, ooff = 0.0, 0, 0
        t0 = time.perf_counter()
        for i, sz in enumerate(sizes):
            used, made = C.c_uint64(0), C.c_uint64(0)
            rc = L.qzb200DecompressDevice(C.byref(sess), d_out + off, h_out + off, sz, d_back + ooff, CALL, C.byref(used), C.byref(made))
            assert rc == 0 and used.value == sz and made.value == CALL, (rc, used.value, sz, made.value)
            kms += prod.stats(sess).kernel_ms; off += sz; ooff += CALL
        wall = time.perf_counter() - t0
    res[name + "_decompress"] = {"GBps_out_wall": round(N / wall / 1e9, 2), "GBps_out_kernels": round(N / (kms / 1e3) / 1e9, 2),
                                 "GBps_in_plus_out_kernels": round((N + total_c) / (kms / 1e3) / 1e9, 2)}
    L.qzb200CopyToHost(h_back, d_back, N)
    assert C.string_at(h_back, 1 << 20) == C.string_at(h_in, 1 << 20) and C.string_at(h_back + N - 4096, 4096) == C.string_at(h_in + N - 4096, 4096)
    # host path decompress (pinned -> pinned)
    for rep in range(2):      # first pass allocates the engine's device/pinned buffers
        t0 = time.perf_counter(); off = ooff = 0
        for sz in sizes:
            rc, used, made = prod.decompress_call(sess, h_out + off, sz, h_back + ooff, CALL)
            assert rc == 0 and made == CALL
            off += sz; ooff += CALL
        res[name + "_decompress"]["GBps_out_e2e_host"] = round(N / (time.perf_counter() - t0) / 1e9, 2)
    prod.end_session(sess)
# the reference's software inflate on all host threads over the same gzip-ext members (CPU baseline of decompress)
if os.path.exists(q.REF_SO) and not os.environ.get("EXTRA_NOCPU") and not ONLY:
    import threading
    ref = q.QzLib(q.REF_SO)
    sess0 = prod.new_session(fmt=q.QZ_DEFLATE_GZIP_EXT)
    rc, used, made, _ = prod.compress_device(sess0, d_in, CALL, d_out, cap, 1)
    L.qzb200CopyToHost(h_out, d_out, made)
    blob = C.string_at(h_out, made)
    prod.end_session(sess0)
    # member boundaries via the QZ extra field
    offs, o = [], 0
    while o < len(blob):
        csz = int.from_bytes(blob[o + 20:o + 24], "little"); offs.append((o, 24 + csz + 8)); o += 24 + csz + 8
    T = os.cpu_count() or 1
    per = (len(offs) + T - 1) // T
    bar = threading.Barrier(T + 1)
    def work(t):
        mine = offs[t * per:(t + 1) * per]
        sess = ref.new_session(fmt=q.QZ_DEFLATE_GZIP_EXT)
        dst = (C.c_ubyte * (65536 * max(1, len(mine))))()
        bar.wait()
        if mine:
            lo, hi = mine[0][0], mine[-1][0] + mine[-1][1]
            rc, used, made = ref.decompress_call(sess, h_out + lo, hi - lo, C.addressof(dst), len(dst))
            assert rc == 0 and made == 65536 * len(mine), (rc, made)
        bar.wait()
        ref.end_session(sess)
    ths = [threading.Thread(target=work, args=(t,)) for t in range(T)]
    [t.start() for t in ths]; bar.wait(); t0 = time.perf_counter(); bar.wait(); dt = time.perf_counter() - t0; [t.join() for t in ths]
    res["cpu_reference_decompress"] = {"GBps_out": round(CALL / dt / 1e9, 2), "threads": T, "sample_MiB_out": CALL >> 20}
# stream API, BASELINE config 5: RAW, 4 KiB submissions (a slice of the 1 GiB stream), driven from C (harness/stream_drive.c)
SN = int(os.environ.get("EXTRA_STREAM_MIB", "64")) << 20
drv = C.CDLL(os.path.join(os.path.dirname(q.CORPUS_SO), "libqzdrive.so"))
drv.qzdrive_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint, C.c_void_p, C.c_uint, C.c_void_p, C.c_size_t,
                               C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint), C.POINTER(C.c_double)]
fn_s, fn_e = C.cast(L.qzCompressStream, C.c_void_p), C.cast(L.qzEndStream, C.c_void_p)
for sb, batch_kb in ((65536, None), (65536, "0"), (2 * 1024 * 1024 - 5 * 1024, "0")) if SN and not ONLY else ():
    if batch_kb is None: os.environ.pop("QZB200_STREAM_BATCH_KB", None)
    else: os.environ["QZB200_STREAM_BATCH_KB"] = batch_kb
    sess = prod.new_session(fmt=q.QZ_DEFLATE_RAW, strm_buff_sz=sb)
    ocap = 8 << 20; obuf = (C.c_ubyte * ocap)()
    outb, calls, crc, secs = C.c_uint64(0), C.c_uint64(0), C.c_uint(0), C.c_double(0)
    rc = drv.qzdrive_stream(fn_s, fn_e, C.byref(sess), h_in, SN, 4096, obuf, ocap, None, 0, C.byref(outb), C.byref(calls), C.byref(crc), C.byref(secs))
    assert rc == 0, rc
    prod.end_session(sess)
    res[f"stream_raw_4KiB_strmbuf_{sb}_{'batched4MiB' if batch_kb is None else 'per_strm_buff'}"] = {"MBps": round(SN / secs.value / 1e6, 1), "calls": calls.value, "ratio": round(outb.value / SN, 4), "stream_MiB": SN >> 20, "driver": "C"}
print(json.dumps(res))
