"""ctypes view of the qatzip.h C ABI (reference include/qatzip.h) -- harness code.

The same binding drives three shared objects, because all three export the same ABI:
  * qatzip_b200/libqatzip.so          the product (sm_100a CUDA behind the C ABI)
  * oracle/_ref/liboracle_qatzip.so   the unmodified reference compiled for the CPU (checker)
and, through OraclePort below, oracle/liboracle_port.so (the C restatement; its own tiny API).

Only tests/, bench.py and __graft_entry__.smoke() import this module.
"""
import ctypes as C
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRODUCT_SO = os.environ.get("QZ_PRODUCT_SO") or os.path.join(ROOT, "qatzip_b200", "libqatzip.so")   # env: A/B builds on the GPU box
REF_SO = os.path.join(ROOT, "oracle", "_ref", "liboracle_qatzip.so")
PORT_SO = os.path.join(ROOT, "oracle", "liboracle_port.so")
CORPUS_SO = os.path.join(ROOT, "harness", "libqzcorpus.so")

# return codes (reference include/qatzip.h:311-361)
QZ_OK, QZ_DUPLICATE, QZ_FORCE_SW = 0, 1, 2
QZ_PARAMS, QZ_FAIL, QZ_BUF_ERROR, QZ_DATA_ERROR, QZ_TIMEOUT = -1, -2, -3, -4, -5
QZ_NO_HW, QZ_NOSW_NO_HW, QZ_NOT_SUPPORTED = 11, -101, -200
# QzDataFormat_T (reference include/qatzip.h:235-245)
QZ_DEFLATE_4B, QZ_DEFLATE_GZIP, QZ_DEFLATE_GZIP_EXT, QZ_DEFLATE_RAW = 0, 1, 2, 3
FMT_LZ4 = 4  # harness-only tag: LZ4 is selected by the session type, not by data_fmt
FMT_ZLIB = 5  # harness-only tag: zlib wire format = qzSetupSessionDeflateExt with zlib_format=1
QZ_DEFLATE, QZ_LZ4 = 8, ord("4")
QZ_DYNAMIC_HDR, QZ_STATIC_HDR = 0, 1
QZ_DIR_COMPRESS, QZ_DIR_DECOMPRESS, QZ_DIR_BOTH = 0, 1, 2
COMMON_MEM, PINNED_MEM = 0, 1

FMT_NAMES = {QZ_DEFLATE_4B: "4B", QZ_DEFLATE_GZIP: "GZIP", QZ_DEFLATE_GZIP_EXT: "GZIP_EXT",
             QZ_DEFLATE_RAW: "RAW", FMT_LZ4: "LZ4", FMT_ZLIB: "ZLIB"}


class QzSession(C.Structure):  # reference include/qatzip.h:676-687
    _fields_ = [("hw_session_stat", C.c_long), ("thd_sess_stat", C.c_int), ("internal", C.c_void_p),
                ("total_in", C.c_ulong), ("total_out", C.c_ulong)]


class QzSessionParams(C.Structure):  # reference include/qatzip.h:461-498
    _fields_ = [("huffman_hdr", C.c_int), ("direction", C.c_int), ("data_fmt", C.c_int),
                ("comp_lvl", C.c_uint), ("comp_algorithm", C.c_ubyte), ("max_forks", C.c_uint),
                ("sw_backup", C.c_ubyte), ("hw_buff_sz", C.c_uint), ("strm_buff_sz", C.c_uint),
                ("input_sz_thrshold", C.c_uint), ("req_cnt_thrshold", C.c_uint),
                ("wait_cnt_thrshold", C.c_uint)]


class QzSessionParamsCommon(C.Structure):  # reference include/qatzip.h:501-538
    _fields_ = [("direction", C.c_int), ("comp_lvl", C.c_uint), ("comp_algorithm", C.c_ubyte),
                ("max_forks", C.c_uint), ("sw_backup", C.c_ubyte), ("hw_buff_sz", C.c_uint),
                ("strm_buff_sz", C.c_uint), ("input_sz_thrshold", C.c_uint),
                ("req_cnt_thrshold", C.c_uint), ("wait_cnt_thrshold", C.c_uint),
                ("polling_mode", C.c_int), ("is_sensitive_mode", C.c_uint)]


class QzSessionParamsDeflate(C.Structure):  # reference include/qatzip.h:541-546
    _fields_ = [("common_params", QzSessionParamsCommon), ("huffman_hdr", C.c_int), ("data_fmt", C.c_int)]


class QzSessionParamsLZ4(C.Structure):  # reference include/qatzip.h:549-551
    _fields_ = [("common_params", QzSessionParamsCommon)]


class QzSessionParamsDeflateExt(C.Structure):  # reference include/qatzip.h:565-569
    _fields_ = [("deflate_params", QzSessionParamsDeflate), ("stop_decompression_stream_end", C.c_ubyte),
                ("zlib_format", C.c_ubyte)]


class QzStream(C.Structure):  # reference include/qatzip.h:2358-2379
    _fields_ = [("in_sz", C.c_uint), ("out_sz", C.c_uint), ("in_", C.c_void_p), ("out", C.c_void_p),
                ("pending_in", C.c_uint), ("pending_out", C.c_uint), ("crc_type", C.c_int),
                ("crc_32", C.c_uint), ("reserved", C.c_ulonglong), ("opaque", C.c_void_p)]


class QzB200Stats(C.Structure):  # include/qatzip_b200.h
    _fields_ = [("kernel_ms", C.c_double), ("codec_ms", C.c_double), ("codec_launches", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("units", C.c_uint64), ("device", C.c_int),
                ("piece_log2", C.c_int), ("hash_bits", C.c_int), ("h2d_ms", C.c_double), ("d2h_ms", C.c_double),
                ("group_blocks", C.c_int), ("devices", C.c_int)]


def _addr(buf):
    """Address of a bytes / bytearray / numpy array / int without copying."""
    if isinstance(buf, int):
        return buf
    if hasattr(buf, "ctypes"):
        return buf.ctypes.data
    if isinstance(buf, (bytes, bytearray, memoryview)):
        return C.addressof(C.c_char.from_buffer(buf)) if isinstance(buf, bytearray) else \
            C.cast(C.c_char_p(bytes(buf)) if not isinstance(buf, bytes) else C.c_char_p(buf), C.c_void_p).value
    raise TypeError(type(buf))


class QzLib:
    """One loaded implementation of the qatzip.h ABI."""

    def __init__(self, path):
        self.path = path
        self.lib = C.CDLL(path, mode=C.RTLD_LOCAL)
        L = self.lib
        P, UP, V = C.POINTER, C.POINTER(C.c_uint), C.c_void_p
        S = P(QzSession)
        L.qzInit.argtypes = [S, C.c_ubyte]
        L.qzSetupSession.argtypes = [S, P(QzSessionParams)]
        L.qzSetupSessionDeflate.argtypes = [S, P(QzSessionParamsDeflate)]
        L.qzSetupSessionDeflateExt.argtypes = [S, P(QzSessionParamsDeflateExt)]
        L.qzSetupSessionLZ4.argtypes = [S, P(QzSessionParamsLZ4)]
        L.qzGetDefaults.argtypes = [P(QzSessionParams)]
        L.qzSetDefaults.argtypes = [P(QzSessionParams)]
        L.qzGetDefaultsDeflate.argtypes = [P(QzSessionParamsDeflate)]
        L.qzSetDefaultsDeflate.argtypes = [P(QzSessionParamsDeflate)]
        L.qzGetDefaultsLZ4.argtypes = [P(QzSessionParamsLZ4)]
        L.qzSetDefaultsLZ4.argtypes = [P(QzSessionParamsLZ4)]
        L.qzGetDefaultsDeflateExt.argtypes = [P(QzSessionParamsDeflateExt)]
        L.qzCompress.argtypes = [S, V, UP, V, UP, C.c_uint]
        L.qzCompressCrc.argtypes = [S, V, UP, V, UP, C.c_uint, P(C.c_ulong)]
        L.qzDecompress.argtypes = [S, V, UP, V, UP]
        L.qzTeardownSession.argtypes = [S]
        L.qzClose.argtypes = [S]
        L.qzMaxCompressedLength.argtypes = [C.c_uint, S]
        L.qzMaxCompressedLength.restype = C.c_uint
        L.qzMalloc.argtypes = [C.c_size_t, C.c_int, C.c_int]
        L.qzMalloc.restype = V
        L.qzFree.argtypes = [V]
        L.qzFree.restype = None
        L.qzMemFindAddr.argtypes = [V]
        L.qzCompressStream.argtypes = [S, P(QzStream), C.c_uint]
        L.qzDecompressStream.argtypes = [S, P(QzStream), C.c_uint]
        L.qzEndStream.argtypes = [S, P(QzStream)]
        L.qzGetDeflateEndOfStream.argtypes = [S, P(C.c_ubyte)]
        L.qzSetLogLevel.argtypes = [C.c_int]
        if hasattr(L, "qzb200DeviceCount"):   # product-only extensions (include/qatzip_b200.h)
            U64P = P(C.c_uint64)
            L.qzb200CompressDevice.argtypes = [S, V, C.c_uint64, V, C.c_uint64, C.c_uint, U64P, U64P, P(C.c_ulong)]
            L.qzb200DecompressDevice.argtypes = [S, V, V, C.c_uint64, V, C.c_uint64, U64P, U64P]
            L.qzb200GetStats.argtypes = [S, P(QzB200Stats)]
            L.qzb200DeviceAlloc.argtypes = [C.c_uint64]
            L.qzb200DeviceAlloc.restype = V
            L.qzb200DeviceFree.argtypes = [V]
            L.qzb200DeviceFree.restype = None
            L.qzb200CopyToDevice.argtypes = [V, V, C.c_uint64]
            L.qzb200CopyToHost.argtypes = [V, V, C.c_uint64]
        if not os.environ.get("QZ_HARNESS_VERBOSE"):
            L.qzSetLogLevel(0)  # LOG_NONE: the CPU reference logs an error per qzInit without QAT hardware

    # ---- sessions -------------------------------------------------------------------------
    def new_session(self, fmt=QZ_DEFLATE_GZIP_EXT, level=1, hw_buff_sz=65536, huffman=QZ_DYNAMIC_HDR,
                    sw_backup=1, strm_buff_sz=65536, input_sz_thrshold=1024, direction=QZ_DIR_BOTH,
                    expect=QZ_OK, init=True, zlib_format=0, stop_at_stream_end=0):
        """qzInit + qzSetupSessionDeflate / qzSetupSessionLZ4 the way reference utils/qzip.c:449-477 does."""
        if fmt == FMT_ZLIB:
            fmt, zlib_format = QZ_DEFLATE_GZIP_EXT, 1
        sess = QzSession()
        if init:
            rc = self.lib.qzInit(C.byref(sess), sw_backup)
            assert rc in (QZ_OK, QZ_DUPLICATE, QZ_NO_HW), f"qzInit rc={rc}"
        if fmt == FMT_LZ4:
            p = QzSessionParamsLZ4()
            self.lib.qzGetDefaultsLZ4(C.byref(p))
            cp = p.common_params
        else:
            p = QzSessionParamsDeflate()
            self.lib.qzGetDefaultsDeflate(C.byref(p))
            p.data_fmt = fmt
            p.huffman_hdr = huffman
            cp = p.common_params
        cp.comp_lvl = level
        cp.hw_buff_sz = hw_buff_sz
        cp.strm_buff_sz = strm_buff_sz
        cp.sw_backup = sw_backup
        cp.input_sz_thrshold = input_sz_thrshold
        cp.direction = direction
        if fmt != FMT_LZ4 and (zlib_format or stop_at_stream_end):
            # reference include/qatzip.h:565-569: the Ext struct selects the zlib wire format / stop-at-stream-end
            px = QzSessionParamsDeflateExt()
            px.deflate_params = p
            px.stop_decompression_stream_end = stop_at_stream_end
            px.zlib_format = zlib_format
            rc = self.lib.qzSetupSessionDeflateExt(C.byref(sess), C.byref(px))
            assert rc == expect, f"qzSetupSessionDeflateExt rc={rc} expected {expect}"
            return sess
        rc = (self.lib.qzSetupSessionLZ4 if fmt == FMT_LZ4 else self.lib.qzSetupSessionDeflate)(C.byref(sess), C.byref(p))
        assert rc == expect, f"qzSetupSession rc={rc} expected {expect}"
        return sess

    def end_session(self, sess):
        self.lib.qzTeardownSession(C.byref(sess))

    def stats(self, sess):
        st = QzB200Stats()
        rc = self.lib.qzb200GetStats(C.byref(sess), C.byref(st))
        assert rc == QZ_OK
        return st

    def compress_device(self, sess, d_src, n, d_dst, cap, last=1):
        used, made, crc = C.c_uint64(0), C.c_uint64(0), C.c_ulong(0)
        rc = self.lib.qzb200CompressDevice(C.byref(sess), d_src, n, d_dst, cap, last, C.byref(used), C.byref(made), C.byref(crc))
        return rc, used.value, made.value, crc.value & 0xFFFFFFFF

    def decompress_device(self, sess, d_src, h_view, n, d_dst, cap):
        used, made = C.c_uint64(0), C.c_uint64(0)
        rc = self.lib.qzb200DecompressDevice(C.byref(sess), d_src, _addr(h_view), n, d_dst, cap, C.byref(used), C.byref(made))
        return rc, used.value, made.value

    # ---- one-shot helpers -----------------------------------------------------------------
    def compress_call(self, sess, src, src_len, dst, dst_cap, last=1, crc=None):
        """Raw call. src/dst: anything _addr() understands. Returns (rc, consumed, produced[, crc])."""
        sl, dl = C.c_uint(src_len), C.c_uint(dst_cap)
        if crc is None:
            rc = self.lib.qzCompress(C.byref(sess), _addr(src), C.byref(sl), _addr(dst), C.byref(dl), last)
            return rc, sl.value, dl.value
        c = C.c_ulong(crc)
        rc = self.lib.qzCompressCrc(C.byref(sess), _addr(src), C.byref(sl), _addr(dst), C.byref(dl), last, C.byref(c))
        return rc, sl.value, dl.value, c.value & 0xFFFFFFFF

    def decompress_call(self, sess, src, src_len, dst, dst_cap):
        sl, dl = C.c_uint(src_len), C.c_uint(dst_cap)
        rc = self.lib.qzDecompress(C.byref(sess), _addr(src), C.byref(sl), _addr(dst), C.byref(dl))
        return rc, sl.value, dl.value

    def compress(self, data, cap=None, **kw):
        """bytes -> bytes through a fresh session; asserts QZ_OK and full consumption."""
        data = bytes(data)
        sess = self.new_session(**kw)
        try:
            cap = cap if cap is not None else len(data) + len(data) // 4 + 4096 * (1 + len(data) // max(1024, kw.get("hw_buff_sz", 65536)))
            dst = bytearray(cap)
            rc, used, made = self.compress_call(sess, data, len(data), dst, cap)
            assert rc == QZ_OK and used == len(data), f"compress rc={rc} used={used}/{len(data)}"
            return bytes(dst[:made])
        finally:
            self.end_session(sess)

    def decompress(self, blob, out_cap, **kw):
        blob = bytes(blob)
        sess = self.new_session(**kw)
        try:
            dst = bytearray(max(out_cap, 1))
            rc, used, made = self.decompress_call(sess, blob, len(blob), dst, out_cap)
            assert rc == QZ_OK and used == len(blob), f"decompress rc={rc} used={used}/{len(blob)} made={made}"
            return bytes(dst[:made])
        finally:
            self.end_session(sess)


class OraclePort:
    """oracle/liboracle_port.so -- the C restatement (checker only)."""

    def __init__(self, path=PORT_SO):
        self.lib = C.CDLL(path)
        L, V, SP = self.lib, C.c_void_p, C.POINTER(C.c_size_t)
        L.qzo_crc32.argtypes = [C.c_uint32, V, C.c_size_t]
        L.qzo_crc32.restype = C.c_uint32
        L.qzo_crc32_combine.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64]
        L.qzo_crc32_combine.restype = C.c_uint32
        L.qzo_xxh32.argtypes = [V, C.c_size_t, C.c_uint32]
        L.qzo_xxh32.restype = C.c_uint32
        L.qzo_adler32.argtypes = [C.c_uint32, V, C.c_size_t]
        L.qzo_adler32.restype = C.c_uint32
        L.qzo_compress.argtypes = [C.c_int, C.c_int, C.c_uint32, V, SP, V, SP, C.c_int, C.POINTER(C.c_uint32)]
        L.qzo_decompress.argtypes = [C.c_int, V, SP, V, SP]
        L.qzo_inflate_raw.argtypes = [V, C.c_size_t, V, C.c_size_t, SP, SP, C.c_int, C.POINTER(C.c_int)]
        L.qzo_lz4_block_compress.argtypes = [V, C.c_size_t, V, C.c_size_t]
        L.qzo_lz4_block_compress.restype = C.c_size_t
        L.qzo_lz4_block_decompress.argtypes = [V, C.c_size_t, V, C.c_size_t, C.c_size_t]
        L.qzo_lz4_block_decompress.restype = C.c_long

    def crc32(self, data, crc=0):
        data = bytes(data)
        return self.lib.qzo_crc32(crc, data, len(data))

    def crc32_combine(self, a, b, len_b):
        return self.lib.qzo_crc32_combine(a, b, len_b)

    def xxh32(self, data, seed=0):
        data = bytes(data)
        return self.lib.qzo_xxh32(data, len(data), seed)

    def adler32(self, data, adler=1):
        data = bytes(data)
        return self.lib.qzo_adler32(adler, data, len(data))

    def compress(self, data, fmt, level=1, hw_buff_sz=65536, last=1, cap=None, want_crc=False):
        data = bytes(data)
        cap = cap if cap is not None else len(data) + len(data) // 4 + 4096 * (1 + len(data) // hw_buff_sz)
        dst = bytearray(cap)
        sl, dl, crc = C.c_size_t(len(data)), C.c_size_t(cap), C.c_uint32(0)
        rc = self.lib.qzo_compress(fmt, level, hw_buff_sz, data, C.byref(sl), _addr(dst), C.byref(dl), last, C.byref(crc))
        assert rc == QZ_OK and sl.value == len(data), f"oracle port compress rc={rc}"
        return (bytes(dst[:dl.value]), crc.value) if want_crc else bytes(dst[:dl.value])

    def decompress_call(self, blob, fmt, out_cap):
        blob = bytes(blob)
        dst = bytearray(max(out_cap, 1))
        sl, dl = C.c_size_t(len(blob)), C.c_size_t(out_cap)
        rc = self.lib.qzo_decompress(fmt, blob, C.byref(sl), _addr(dst), C.byref(dl))
        return rc, sl.value, bytes(dst[:dl.value])

    def decompress(self, blob, fmt, out_cap):
        rc, used, out = self.decompress_call(blob, fmt, out_cap)
        assert rc == QZ_OK and used == len(blob), f"oracle port decompress rc={rc} used={used}/{len(blob)}"
        return out


class Corpus:
    """harness/libqzcorpus.so -- deterministic synthetic inputs."""
    REF_RLE, SILESIA_LIKE = 0, 1
    SILESIA_SEED = 0x51CE51A

    def __init__(self, path=CORPUS_SO):
        self.lib = C.CDLL(path)
        self.lib.qzcorpus_fill.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_void_p, C.c_size_t, C.c_int]

    def fill(self, kind, addr, nbytes, seed=None, first_seg=0, threads=None):
        seed = (self.SILESIA_SEED if kind == self.SILESIA_LIKE else 1) if seed is None else seed
        threads = threads or min(32, os.cpu_count() or 1)
        rc = self.lib.qzcorpus_fill(kind, seed, first_seg, addr, nbytes, threads)
        assert rc == 0

    def make(self, kind, nbytes, **kw):
        buf = bytearray(nbytes)
        if nbytes:
            self.fill(kind, _addr(buf), nbytes, **kw)
        return bytes(buf)
