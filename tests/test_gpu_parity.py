"""GPU parity tests: everything goes through the C ABI of qatzip_b200/libqatzip.so and is judged by
the oracle (oracle/_ref = the reference's own software path when built, and the C restatement).

Matrix:
  1. ours.compress -> oracle.decompress == input           (every format x sizes x hw_buff_sz)
  2. oracle.compress -> ours.decompress == input           (single big members, multi-block zlib output)
  3. ours <-> ours round trips incl. the stream API slice patterns of reference test mode 9
  4. qzCompressCrc / strm.crc_32 == zlib crc32; LZ4 footer == XXH32; gzip footer CRC/ISIZE; QZ extra sizes
  5. error semantics of reference test modes 17 / 13-16 / 22
  6. ratio within 5 % of the oracle at level 1
and, at sizes the CPU cannot check byte by byte in seconds, size-independent properties
(compress -> decompress identity on the GPU, CRC of CRCs)."""
import ctypes as C
import os
import zlib

import pytest

from harness import qzapi as q
from conftest import has_gpu

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not has_gpu(), reason="needs a CUDA device")]

DEFLATE_FMTS = [q.QZ_DEFLATE_4B, q.QZ_DEFLATE_GZIP, q.QZ_DEFLATE_GZIP_EXT, q.QZ_DEFLATE_RAW]
ALL_FMTS = DEFLATE_FMTS + [q.FMT_LZ4, q.FMT_ZLIB]
SIZES = [0, 1, 127, 1023, 1024, 4096, 8191, 8192, 8193, 65535, 65536, 65537, 524288, 1060921]


@pytest.fixture(scope="module")
def data(corpus):
    return corpus.make(q.Corpus.SILESIA_LIKE, 12 << 20)


def pick(data, n, k=0):
    off = ((k * 2654435761) % (len(data) - n - 1)) if n < len(data) else 0
    return data[off:off + n]


@pytest.mark.parametrize("fmt", ALL_FMTS)
def test_ours_to_oracle_all_sizes(prod, port, ref, data, fmt):
    for i, n in enumerate(SIZES):
        d = pick(data, n, i + fmt)
        blob = prod.compress(d, fmt=fmt)
        assert port.decompress(blob, fmt, n + 8) == d, f"oracle port could not restore n={n}"
        if n:
            assert ref.decompress(blob, n + 8, fmt=fmt) == d, f"reference software path could not restore n={n}"
    assert len(prod.compress(b"", fmt=q.QZ_DEFLATE_GZIP_EXT)) == 34       # QZ_COMPRESSED_SZ_OF_EMPTY_FILE


@pytest.mark.parametrize("hw", [1024, 4096, 65536, 524288])
@pytest.mark.parametrize("fmt", ALL_FMTS)
def test_ours_to_oracle_chunk_sizes(prod, port, ref, data, fmt, hw):
    """reference bt.c sweep + mode 17 (4 KiB input with hw_buff_sz 1 KiB)"""
    for n in (4096, hw - 1, hw, hw + 1, 3 * hw + 17, 600000):
        d = pick(data, n, hw + n)
        blob = prod.compress(d, fmt=fmt, hw_buff_sz=hw)
        assert port.decompress(blob, fmt, n + 8) == d
        assert ref.decompress(blob, n + 8, fmt=fmt, hw_buff_sz=hw) == d
        assert prod.decompress(blob, n + 8, fmt=fmt, hw_buff_sz=hw) == d


@pytest.mark.parametrize("fmt", ALL_FMTS)
def test_oracle_to_ours(prod, port, ref, data, fmt):
    for i, n in enumerate(SIZES[1:]):
        d = pick(data, n, 3 * i + fmt)
        assert prod.decompress(port.compress(d, fmt), n + 8, fmt=fmt) == d, f"port-made stream n={n}"
        # the reference software path emits ONE member / frame per call, with Z_FULL_FLUSH blocks inside
        assert prod.decompress(ref.compress(d, fmt=fmt), n + 8, fmt=fmt) == d, f"reference-made stream n={n}"


def test_oracle_levels_and_strategies(prod, data):
    """SW L9 compress -> our decompress (reference mode 17), fixed and stored blocks from zlib"""
    d = pick(data, 4 << 20, 5)
    for level, strat in ((9, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_FIXED), (0, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_HUFFMAN_ONLY)):
        co = zlib.compressobj(level, zlib.DEFLATED, 31, 9, strat)
        blob = co.compress(d) + co.flush()
        assert prod.decompress(blob, len(d) + 8, fmt=q.QZ_DEFLATE_GZIP) == d


def test_mixed_members(prod, ref, port, data):
    """reference test mode 5: alternate hardware-format members and software members in one buffer"""
    parts, blob = [], b""
    for i, n in enumerate((70000, 5000, 200000, 1, 65536)):
        d = pick(data, n, 11 * i)
        parts.append(d)
        blob += prod.compress(d, fmt=q.QZ_DEFLATE_GZIP_EXT) if i % 2 == 0 else ref.compress(d, fmt=q.QZ_DEFLATE_GZIP)
    whole = b"".join(parts)
    assert prod.decompress(blob, len(whole) + 8, fmt=q.QZ_DEFLATE_GZIP_EXT) == whole
    assert prod.decompress(blob, len(whole) + 8, fmt=q.QZ_DEFLATE_GZIP) == whole


@pytest.mark.parametrize("fmt", DEFLATE_FMTS)
def test_compress_crc_equals_zlib(prod, data, fmt):
    """reference test/main.c:4283-4337 plus multi-chunk and multi-call accumulation"""
    for n in (65536, 1023, 300000, 5 << 20):
        d = pick(data, n, n)
        sess = prod.new_session(fmt=fmt)
        dst = bytearray(n + n // 4 + 65536)
        rc, used, made, crc = prod.compress_call(sess, d, n, dst, len(dst), crc=0)
        assert rc == q.QZ_OK and used == n and crc == zlib.crc32(d)
        # second call continues the running value
        rc, used, made, crc2 = prod.compress_call(sess, d[:1000], 1000, dst, len(dst), crc=crc)
        assert crc2 == zlib.crc32(d + d[:1000])
        prod.end_session(sess)


def test_footers_and_extra_field(prod, port, data):
    d = pick(data, 200000, 9)
    ext = prod.compress(d, fmt=q.QZ_DEFLATE_GZIP_EXT)
    off, pos = 0, 0
    while off < len(ext):
        assert ext[off:off + 16] == bytes([0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 12, 0, ord("Q"), ord("Z"), 8, 0])
        s, c = int.from_bytes(ext[off + 16:off + 20], "little"), int.from_bytes(ext[off + 20:off + 24], "little")
        assert zlib.decompress(ext[off + 24:off + 24 + c], -15) == d[pos:pos + s]
        f = ext[off + 24 + c:off + 32 + c]
        assert int.from_bytes(f[:4], "little") == zlib.crc32(d[pos:pos + s]) and int.from_bytes(f[4:], "little") == s
        off += 32 + c; pos += s
    assert pos == len(d)
    lz = prod.compress(d, fmt=q.FMT_LZ4)
    assert lz[:6] == bytes([4, 0x22, 0x4d, 0x18, 0x4c, 0x40]) and int.from_bytes(lz[6:14], "little") == 65536
    assert lz[14] == (port.xxh32(lz[4:14]) >> 8) & 0xff
    assert int.from_bytes(lz[-4:], "little") == port.xxh32(d[196608:]) and lz[-8:-4] == bytes(4)
    b4 = prod.compress(d, fmt=q.QZ_DEFLATE_4B)
    n0 = int.from_bytes(b4[:4], "little")
    assert zlib.decompress(b4[4:4 + n0], -15) == d[:65536]


def test_static_huffman_session(prod, port, data):
    d = pick(data, 300000, 4)
    dyn = prod.compress(d, fmt=q.QZ_DEFLATE_GZIP_EXT)
    fix = prod.compress(d, fmt=q.QZ_DEFLATE_GZIP_EXT, huffman=q.QZ_STATIC_HDR)
    assert port.decompress(fix, q.QZ_DEFLATE_GZIP_EXT, len(d) + 8) == d and len(fix) > len(dyn)
    assert (fix[24] >> 1) & 3 in (0, 1)        # first block is stored or fixed, never dynamic


def ratio_corpora(corpus):
    """the three corpora of the ratio gates: SILESIA-LIKE (the bench workload), REF-RLE (the distribution of the reference's own
    test generator, reference test/main.c:293-310) and real text (tests/golden/text_sample.txt: this repository's prose and code)"""
    here = os.path.dirname(os.path.abspath(__file__))
    return [("SILESIA-LIKE", corpus.make(q.Corpus.SILESIA_LIKE, 24 << 20)), ("REF-RLE", corpus.make(q.Corpus.REF_RLE, 4 << 20)),
            ("real text", open(os.path.join(here, "golden", "text_sample.txt"), "rb").read())]


def test_ratio_within_5_percent_of_reference(prod, ref, corpus):
    """the BASELINE gate on every corpus: len(ours) / len(reference zlib -1, same hw_buff_sz) <= 1.05"""
    for name, d in ratio_corpora(corpus):
        ours = len(prod.compress(d, fmt=q.QZ_DEFLATE_GZIP_EXT))
        theirs = len(ref.compress(d, fmt=q.QZ_DEFLATE_GZIP_EXT))
        print(f"deflate {name}: ours {ours / len(d):.4f} reference {theirs / len(d):.4f} rel {ours / theirs - 1:+.2%}")
        assert ours <= theirs * 1.05, name


def test_compression_levels(prod, ref):
    """comp_lvl is honoured the way QAT 2.0 does it (reference README.md:133-148: levels 1-5, 6-8 and 9-12 are three search depths):
    levels 6 and up search two entries per hash bucket.  Every level decodes with the reference; on text the deeper search
    is smaller."""
    here = os.path.dirname(os.path.abspath(__file__))
    d = open(os.path.join(here, "golden", "text_sample.txt"), "rb").read()
    sizes = {}
    for lvl in (1, 5, 6, 9):
        blob = prod.compress(d, fmt=q.QZ_DEFLATE_GZIP_EXT, level=lvl)
        assert ref.decompress(blob, len(d) + 8, fmt=q.QZ_DEFLATE_GZIP_EXT) == d
        sizes[lvl] = len(blob)
    print("levels", sizes)
    assert sizes[1] == sizes[5] and sizes[6] == sizes[9] and sizes[6] < sizes[1]
    # LZ4: levels below 6 use sixteen matchers with smaller tables, 6 and up twelve with larger ones
    lz = {}
    for lvl in (1, 9):
        blob = prod.compress(d, fmt=q.FMT_LZ4, level=lvl)
        assert ref.decompress(blob, len(d) + 8, fmt=q.FMT_LZ4) == d
        lz[lvl] = len(blob)
    print("LZ4 levels", lz)
    assert lz[9] < lz[1]


def test_lz4_ratio_within_5_percent_of_reference(prod, ref, corpus):
    """LZ4 at the same granularity as the hardware path: one frame with one 64 KiB block per chunk (reference session:
    lz4BlockMaxSize = 64 KiB, src/qatzip_utils.c:292-298), against the reference's LZ4F per 64 KiB chunk: <= 1.05"""
    for name, d in ratio_corpora(corpus):
        d = d[:8 << 20]
        ours = len(prod.compress(d, fmt=q.FMT_LZ4))
        theirs = sum(len(ref.compress(d[i:i + 65536], fmt=q.FMT_LZ4)) for i in range(0, len(d), 65536))
        print(f"LZ4 {name}: ours {ours / len(d):.4f} reference {theirs / len(d):.4f} rel {ours / theirs - 1:+.2%}")
        assert ours <= theirs * 1.05, name


# ---------------------------------------------------------------------------- error semantics
def test_buf_error_partial_progress_compress(prod, port, data):
    """reference mode 17 (:4212-4271): small dest -> QZ_BUF_ERROR, whole chunks only"""
    d = pick(data, 300000, 1)
    sess = prod.new_session(fmt=q.QZ_DEFLATE_GZIP_EXT)
    dst = bytearray(1024)
    rc, used, made = prod.compress_call(sess, d, len(d), dst, 1024)
    assert rc == q.QZ_BUF_ERROR and used == 0 and made == 0
    full = prod.compress(d, fmt=q.QZ_DEFLATE_GZIP_EXT)
    cap = len(full) - 10                                   # everything but the last member fits
    dst = bytearray(cap)
    rc, used, made = prod.compress_call(sess, d, len(d), dst, cap)
    assert rc == q.QZ_BUF_ERROR and used == 4 * 65536 and 0 < made <= cap
    assert port.decompress(bytes(dst[:made]), q.QZ_DEFLATE_GZIP_EXT, used + 8) == d[:used]
    prod.end_session(sess)


def test_decompress_errors(prod, data):
    d = pick(data, 200000, 2)
    blob = prod.compress(d, fmt=q.QZ_DEFLATE_GZIP_EXT)
    sess = prod.new_session(fmt=q.QZ_DEFLATE_GZIP_EXT)
    out = bytearray(len(d) + 8)
    bad = bytes([0x1e]) + blob[1:]                                        # corrupt id1 -> QZ_FAIL (:3858-3866)
    rc, used, made = prod.decompress_call(sess, bad, len(bad), out, len(out))
    assert rc == q.QZ_FAIL and used == 0 and made == 0
    garb = bytearray(blob); garb[60:100] = bytes(40)                      # garbage payload
    rc, used, made = prod.decompress_call(sess, bytes(garb), len(garb), out, len(out))
    assert rc in (q.QZ_FAIL, q.QZ_DATA_ERROR) and used == 0
    flip = bytearray(blob); flip[-6] ^= 0x55                              # CRC of the last member
    rc, used, made = prod.decompress_call(sess, bytes(flip), len(flip), out, len(out))
    assert rc == q.QZ_DATA_ERROR and made == 3 * 65536 and bytes(out[:made]) == d[:made]
    rc, used, made = prod.decompress_call(sess, blob, len(blob), out, 1024)        # 1 KiB dest -> QZ_BUF_ERROR
    assert rc == q.QZ_BUF_ERROR and used == 0 and made == 0
    rc, used, made = prod.decompress_call(sess, blob, len(blob), out, 70000)       # room for one member only
    assert rc == q.QZ_BUF_ERROR and made == 65536 and bytes(out[:made]) == d[:65536]
    rc, used, made = prod.decompress_call(sess, blob[:-5], len(blob) - 5, out, len(out))   # truncated tail
    assert rc == q.QZ_DATA_ERROR and made == 3 * 65536
    rc, used, made = prod.decompress_call(sess, blob, 0, out, len(out))            # *src_len == 0 -> QZ_OK (:2465)
    assert rc == q.QZ_OK and made == 0
    prod.end_session(sess)


def test_decompress_with_other_hw_buff_sz(prod, data):
    """reference mode 7 / mode 17 :4039: decode with a smaller or larger hw_buff_sz than the encoder's"""
    d = pick(data, 400000, 6)
    blob = prod.compress(d, fmt=q.QZ_DEFLATE_GZIP_EXT, hw_buff_sz=65536)
    for hw in (32768, 65536, 131072):
        assert prod.decompress(blob, len(d) + 8, fmt=q.QZ_DEFLATE_GZIP_EXT, hw_buff_sz=hw) == d


def test_stop_at_stream_end(prod, data):
    """reference modes 27/30: stop_decompression_stream_end decodes exactly one member"""
    d = pick(data, 512 * 1024, 8)
    blob = prod.compress(d, fmt=q.QZ_DEFLATE_GZIP_EXT)
    sess = q.QzSession()
    p = q.QzSessionParamsDeflateExt()
    prod.lib.qzGetDefaultsDeflateExt(C.byref(p))
    p.stop_decompression_stream_end = 1
    assert prod.lib.qzSetupSessionDeflateExt(C.byref(sess), C.byref(p)) == q.QZ_OK
    out = bytearray(len(d))
    rc, used, made = prod.decompress_call(sess, blob, len(blob), out, len(out))
    assert rc == q.QZ_OK and made == 65536 and bytes(out[:made]) == d[:65536]
    eos = C.c_ubyte(0)
    assert prod.lib.qzGetDeflateEndOfStream(C.byref(sess), C.byref(eos)) == q.QZ_OK and eos.value == 1
    prod.end_session(sess)


def test_zlib_format(prod, port, ref, data, monkeypatch):
    """zlib_format=1 sessions (reference include/qatzip.h:565-569, src/qatzip_gzip.c:263-306): one RFC 1950
    stream per chunk, readable by any zlib; decode finds the streams without size fields."""
    d = pick(data, 3 * 1024 * 1024 + 777, 4)
    blob = prod.compress(d, fmt=q.FMT_ZLIB)
    out, rest, n = b"", blob, 0
    while rest:
        assert rest[:2] == b"\x78\x9c"
        z = zlib.decompressobj()
        out += z.decompress(rest)
        assert z.eof
        rest, n = z.unused_data, n + 1
    assert out == d and n == (len(d) + 65535) // 65536
    # running checksum of a zlib session is the Adler-32 of everything consumed
    sess = prod.new_session(fmt=q.FMT_ZLIB)
    dst = bytearray(len(d))
    rc, used, made, ck = prod.compress_call(sess, d, len(d), dst, len(dst), crc=0)
    assert rc == q.QZ_OK and ck == zlib.adler32(d)
    prod.end_session(sess)
    # streams made by zlib itself at other levels / window sizes, concatenated
    parts = [pick(data, n, 7 * i) for i, n in enumerate((100000, 1, 70000, 300000))]
    cat = b"".join(zlib.compress(p, lvl) for p, lvl in zip(parts, (1, 6, 9, 0)))
    co = zlib.compressobj(6, zlib.DEFLATED, 9)
    small = pick(data, 50000, 3)
    cat += co.compress(small) + co.flush()
    parts.append(small)
    assert prod.decompress(cat, sum(map(len, parts)) + 8, fmt=q.FMT_ZLIB) == b"".join(parts)
    # payload full of bytes that look like zlib headers: stored blocks of 78 9C 78 01 78 DA ...
    noisy = (b"\x78\x9c\x78\x01\x78\xda\x78\x5e" * 8 + os.urandom(64)) * 4096
    nb = prod.compress(noisy, fmt=q.FMT_ZLIB)
    assert prod.decompress(nb, len(noisy) + 8, fmt=q.FMT_ZLIB) == noisy
    assert ref.decompress(nb, len(noisy) + 8, fmt=q.FMT_ZLIB) == noisy
    # several discovery windows, a stream larger than one window, and streams cut by a window end
    monkeypatch.setenv("QZB200_ZLIB_WINDOW_MB", "1")
    big = pick(data, 6 * 1024 * 1024, 1)
    bb = prod.compress(big, fmt=q.FMT_ZLIB)
    assert len(bb) > 2 << 20
    assert prod.decompress(bb, len(big) + 8, fmt=q.FMT_ZLIB) == big
    one = zlib.compress(big[:3 << 20], 0)             # one stream of stored blocks, three windows long
    assert prod.decompress(bb[:len(bb)] + one + bb, 3 * len(big), fmt=q.FMT_ZLIB) == big + big[:3 << 20] + big
    monkeypatch.delenv("QZB200_ZLIB_WINDOW_MB")
    # errors: corrupted trailer, corrupted header, truncated tail (whole streams before it are delivered)
    bad = bytearray(blob); bad[-1] ^= 0x55
    sess = prod.new_session(fmt=q.FMT_ZLIB)
    outb = bytearray(len(d) + 8)
    rc, used, made = prod.decompress_call(sess, bytes(bad), len(bad), outb, len(outb))
    assert rc == q.QZ_DATA_ERROR and made == (n - 1) * 65536 and bytes(outb[:made]) == d[:made]
    bad = bytearray(blob); bad[0] = 0x79
    rc, used, made = prod.decompress_call(sess, bytes(bad), len(bad), outb, len(outb))
    assert rc == q.QZ_FAIL and used == 0 and made == 0
    rc, used, made = prod.decompress_call(sess, blob[:-3], len(blob) - 3, outb, len(outb))
    assert rc == q.QZ_DATA_ERROR and made == (n - 1) * 65536 and bytes(outb[:made]) == d[:made]
    # too little room: whole streams that fit are delivered
    rc, used, made = prod.decompress_call(sess, blob, len(blob), outb, 200000)
    assert rc == q.QZ_BUF_ERROR and made == 3 * 65536 and bytes(outb[:made]) == d[:made]
    prod.end_session(sess)
    # stop_decompression_stream_end: one stream, end-of-stream flag set
    sess = prod.new_session(fmt=q.FMT_ZLIB, stop_at_stream_end=1)
    rc, used, made = prod.decompress_call(sess, blob, len(blob), outb, len(outb))
    eos = C.c_ubyte(0)
    assert rc == q.QZ_OK and made == 65536 and bytes(outb[:made]) == d[:65536]
    assert prod.lib.qzGetDeflateEndOfStream(C.byref(sess), C.byref(eos)) == q.QZ_OK and eos.value == 1
    assert blob[used:used + 2] == b"\x78\x9c"
    prod.end_session(sess)


# ---------------------------------------------------------------------------- memory + device entry points
def test_qzmalloc_pinned_zero_copy(prod, port, data):
    L = prod.lib
    n = 3 << 20
    d = pick(data, n, 3)
    for _ in range(50):                                     # reference mode 2: repeated pinned + common allocations
        a, b = L.qzMalloc(100 * 1024, 0, q.PINNED_MEM), L.qzMalloc(100 * 1024, 0, q.COMMON_MEM)
        assert a and b and L.qzMemFindAddr(a) == 1 and L.qzMemFindAddr(b) == 1
        L.qzFree(a); L.qzFree(b)
    src, cap = L.qzMalloc(n, 0, q.PINNED_MEM), n + n // 4 + 65536
    dst = L.qzMalloc(cap, 0, q.PINNED_MEM)
    C.memmove(src, d, n)
    sess = prod.new_session(fmt=q.QZ_DEFLATE_GZIP_EXT)
    rc, used, made = prod.compress_call(sess, src, n, dst, cap)
    assert rc == q.QZ_OK and used == n
    blob = C.string_at(dst, made)
    assert port.decompress(blob, q.QZ_DEFLATE_GZIP_EXT, n + 8) == d
    C.memmove(dst, blob, made)
    rc, used, made2 = prod.decompress_call(sess, dst, made, src, n)
    assert rc == q.QZ_OK and made2 == n and C.string_at(src, n) == d
    prod.end_session(sess); L.qzFree(src); L.qzFree(dst)


@pytest.mark.parametrize("fmt", [q.QZ_DEFLATE_GZIP_EXT, q.QZ_DEFLATE_GZIP, q.FMT_LZ4])
def test_device_resident_entry_points(prod, port, data, fmt):
    L = prod.lib
    n = 8 << 20
    d = data[:n]
    sess = prod.new_session(fmt=fmt)
    cap = n + n // 4 + 65536
    d_in, d_out, d_back = L.qzb200DeviceAlloc(n), L.qzb200DeviceAlloc(cap), L.qzb200DeviceAlloc(n)
    assert d_in and d_out and d_back and L.qzb200CopyToDevice(d_in, d, n) == 0
    rc, used, made, crc = prod.compress_device(sess, d_in, n, d_out, cap)
    assert rc == q.QZ_OK and used == n
    if fmt != q.FMT_LZ4:
        assert crc == zlib.crc32(d)
    st = prod.stats(sess)
    assert st.kernel_launches >= 4 and st.codec_ms > 0
    blob = bytearray(made)
    assert L.qzb200CopyToHost(q._addr(blob), d_out, made) == 0
    assert port.decompress(bytes(blob), fmt, n + 8) == d
    rc, used2, made2 = prod.decompress_device(sess, d_out, blob, made, d_back, n)
    assert rc == q.QZ_OK and used2 == made and made2 == n
    back = bytearray(n)
    L.qzb200CopyToHost(q._addr(back), d_back, n)
    assert bytes(back) == d
    rc, used, made3, _ = prod.compress_device(sess, d_in, n, d_out, 100000)     # dest too small: whole chunks only
    assert rc == q.QZ_BUF_ERROR and used % 65536 == 0 and used < n and made3 <= 100000
    for p in (d_in, d_out, d_back):
        L.qzb200DeviceFree(p)
    prod.end_session(sess)


@pytest.mark.parametrize("fmt", [q.QZ_DEFLATE_GZIP_EXT, q.FMT_LZ4])
def test_device_resident_unaligned_source(prod, port, data, fmt):
    """the window kernels copy a window in by one TMA bulk copy when its address is 16-byte aligned; any other source (here:
    device pointer + 3, ragged length) is copied by hand.  Same stream either way."""
    L = prod.lib
    n = (2 << 20) + 4321
    d = data[:n]
    sess = prod.new_session(fmt=fmt)
    cap = n + n // 4 + 65536
    d_in, d_out = L.qzb200DeviceAlloc(n + 16), L.qzb200DeviceAlloc(cap)
    assert d_in and d_out
    blobs = []
    for skew in (0, 3):
        assert L.qzb200CopyToDevice(d_in + skew, d, n) == 0
        rc, used, made, crc = prod.compress_device(sess, d_in + skew, n, d_out, cap)
        assert rc == q.QZ_OK and used == n
        blob = bytearray(made)
        assert L.qzb200CopyToHost(q._addr(blob), d_out, made) == 0
        assert port.decompress(bytes(blob), fmt, n + 8) == d
        blobs.append(bytes(blob))
    assert blobs[0] == blobs[1]
    L.qzb200DeviceFree(d_in); L.qzb200DeviceFree(d_out)
    prod.end_session(sess)


def test_large_roundtrip_properties(prod, corpus):
    """256 MiB per call (BASELINE-size pieces): GPU compress -> GPU decompress identity and CRC of CRCs."""
    L = prod.lib
    n = 256 << 20
    src = L.qzMalloc(n, 0, q.PINNED_MEM)
    back = L.qzMalloc(n, 0, q.PINNED_MEM)
    cap = L.qzMaxCompressedLength(n, None)
    dst = L.qzMalloc(cap, 0, q.PINNED_MEM)
    corpus.fill(q.Corpus.SILESIA_LIKE, src, n, first_seg=100)
    for fmt in (q.QZ_DEFLATE_GZIP_EXT, q.FMT_LZ4):
        sess = prod.new_session(fmt=fmt)
        rc, used, made, crc = prod.compress_call(sess, src, n, dst, cap, crc=0)
        assert rc == q.QZ_OK and used == n
        if fmt != q.FMT_LZ4:
            assert crc == zlib.crc32(C.string_at(src, n))
        rc, used2, made2 = prod.decompress_call(sess, dst, made, back, n)
        assert rc == q.QZ_OK and used2 == made and made2 == n
        assert zlib.crc32(C.string_at(back, n)) == zlib.crc32(C.string_at(src, n))
        print(f"{q.FMT_NAMES[fmt]} 256 MiB ratio {made / n:.4f}")
        prod.end_session(sess)
    for p in (src, back, dst):
        L.qzFree(p)


# ---------------------------------------------------------------------------- stream API (reference mode 9-12, 20, 22)
def stream_compress(prod, sess, d, slice_sz, out_sz=1 << 20):
    st = q.QzStream()
    out = bytearray()
    obuf = (C.c_ubyte * out_sz)()
    src = (C.c_ubyte * max(len(d), 1)).from_buffer_copy(d if d else b"\0")
    consumed, calls, empties = 0, 0, 0
    while True:
        left = len(d) - consumed
        n = min(slice_sz, left)
        last = 1 if left - n == 0 else 0
        st.in_ = C.addressof(src) + consumed; st.in_sz = n; st.out = C.addressof(obuf); st.out_sz = out_sz
        rc = prod.lib.qzCompressStream(C.byref(sess), C.byref(st), last)
        assert rc == q.QZ_OK, rc
        consumed += st.in_sz; out += bytes(obuf[:st.out_sz]); calls += 1; empties += st.out_sz == 0
        if last and st.pending_in == 0 and st.pending_out == 0 and consumed == len(d):
            break
        assert calls < 10_000_000
    crc = st.crc_32
    prod.lib.qzEndStream(C.byref(sess), C.byref(st))
    return bytes(out), crc, calls, empties


def stream_decompress(prod, sess, blob, slice_sz, out_sz):
    st = q.QzStream()
    out = bytearray()
    obuf = (C.c_ubyte * out_sz)()
    src = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
    consumed, calls = 0, 0
    while True:
        left = len(blob) - consumed
        n = min(slice_sz, left)
        last = 1 if left - n == 0 else 0
        st.in_ = C.addressof(src) + consumed; st.in_sz = n; st.out = C.addressof(obuf); st.out_sz = out_sz
        rc = prod.lib.qzDecompressStream(C.byref(sess), C.byref(st), last)
        assert rc == q.QZ_OK, rc
        consumed += st.in_sz; out += bytes(obuf[:st.out_sz]); calls += 1
        if consumed == len(blob) and st.pending_in == 0 and st.pending_out == 0:
            break
        assert calls < 1_000_000
    prod.lib.qzEndStream(C.byref(sess), C.byref(st))
    return bytes(out)


@pytest.mark.parametrize("batch_kb", [None, "0"])
@pytest.mark.parametrize("fmt", [q.QZ_DEFLATE_GZIP_EXT, q.QZ_DEFLATE_RAW])
def test_stream_compress_slices(prod, port, ref, data, fmt, batch_kb, monkeypatch):
    """batch_kb None: default staging (4 MiB of chunks per engine call); "0": the reference's cadence,
    one engine call per strm_buff_sz (src/qatzip_stream.c:514-560).  Same chunks, same bytes either way."""
    if batch_kb is not None:
        monkeypatch.setenv("QZB200_STREAM_BATCH_KB", batch_kb)
    d = pick(data, 1 << 20, 12)
    sess = prod.new_session(fmt=fmt)
    blobs = []
    for slice_sz in (16384, 4096, 65536, 100000):            # hw_buff_sz/4 like mode 9, 4 KiB like BASELINE config 5
        blob, crc, calls, empties = stream_compress(prod, sess, d, slice_sz)
        assert crc == zlib.crc32(d)                          # strm.crc_32 = CRC-32 of all stream input (mode 11)
        assert port.decompress(blob, fmt, len(d) + 8) == d
        assert ref.decompress(blob, len(d) + 8, fmt=fmt) == d
        blobs.append(blob)
        if slice_sz == 4096:
            # reference cadence: batched to strm_buff_sz, 16 flushes; default: one flush
            assert calls == 256 and empties == (240 if batch_kb == "0" else 255)
    assert all(b == blobs[0] for b in blobs)                 # the stream does not depend on how the input was sliced
    # RAW stream wrapped by hand in a gzip header + {crc_32, in_sz} trailer (mode 11 :3045-3072)
    if fmt == q.QZ_DEFLATE_RAW:
        blob, crc, _, _ = stream_compress(prod, sess, d, 16384)
        gz = bytes([0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 3]) + blob + crc.to_bytes(4, "little") + len(d).to_bytes(4, "little")
        assert zlib.decompress(gz, 31) == d
    prod.end_session(sess)


@pytest.mark.parametrize("workers", ["1", "2"])
@pytest.mark.parametrize("fmt", [q.QZ_DEFLATE_GZIP_EXT, q.QZ_DEFLATE_RAW])
def test_stream_many_jobs_in_flight(prod, port, data, fmt, workers, monkeypatch):
    """Staging of 256 KiB (four chunks a job) over 6 MiB: two dozen jobs go through the stream's worker threads, with an
    output buffer small enough that finished jobs wait for room while the next ones are already being compressed.
    Output order, the combined CRC-32 and the bytes must not depend on the number of workers or on the buffer sizes."""
    monkeypatch.setenv("QZB200_STREAM_BATCH_KB", "256")
    monkeypatch.setenv("QZB200_STREAM_WORKERS", workers)
    d = pick(data, 6 << 20, 41)
    sess = prod.new_session(fmt=fmt)
    blobs = []
    for slice_sz, out_sz in ((50000, 8192), (4096, 1 << 20), (700000, 300000)):
        blob, crc, _, _ = stream_compress(prod, sess, d, slice_sz, out_sz=out_sz)
        assert crc == zlib.crc32(d)
        blobs.append(blob)
    assert all(b == blobs[0] for b in blobs)
    assert port.decompress(blobs[0], fmt, len(d) + 8) == d
    assert blobs[0] == prod.compress(d, fmt=fmt)              # the same chunks as one qzCompress call makes
    prod.end_session(sess)


def test_stream_small_output_pending(prod, port, data):
    """mode 20: 8 KiB out_sz with pending_out draining"""
    d = pick(data, 300000, 13)
    sess = prod.new_session(fmt=q.QZ_DEFLATE_GZIP_EXT)
    blob, crc, _, _ = stream_compress(prod, sess, d, 16384, out_sz=8192)
    assert port.decompress(blob, q.QZ_DEFLATE_GZIP_EXT, len(d) + 8) == d and crc == zlib.crc32(d)
    prod.end_session(sess)


def test_stream_decompress_slices(prod, data):
    d = pick(data, 1 << 20, 14)
    blob = prod.compress(d, fmt=q.QZ_DEFLATE_GZIP_EXT)
    sess = prod.new_session(fmt=q.QZ_DEFLATE_GZIP_EXT)
    for slice_sz, out_sz in ((16384, 1 << 20), (256, 65536), (len(blob), 8192), (70000, 300000)):
        assert stream_decompress(prod, sess, blob, slice_sz, out_sz) == d
    prod.end_session(sess)
    # incompressible members are larger than strm_buff_sz: staging must grow, not fail
    r = os.urandom(200000)
    blob = prod.compress(r, fmt=q.QZ_DEFLATE_GZIP_EXT)
    sess = prod.new_session(fmt=q.QZ_DEFLATE_GZIP_EXT)
    assert stream_decompress(prod, sess, blob, 5000, 1 << 20) == r
    prod.end_session(sess)


def test_golden_streams_through_the_gpu_decoder(prod):
    """the committed fixtures (streams the compiled reference made, tests/golden/make_golden.py) through the product's
    qzDecompress at the hw_buff_sz they were made with, and through qzDecompressStream in 100-byte slices"""
    import hashlib
    import json
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    man = json.load(open(os.path.join(gold, "manifest.json")))
    assert man["cases"]
    for case in man["cases"]:
        raw = open(os.path.join(gold, case["input"]), "rb").read()
        blob = open(os.path.join(gold, case["stream"]), "rb").read()
        assert hashlib.sha256(raw).hexdigest() == case["input_sha256"]
        assert prod.decompress(blob, len(raw) + 8, fmt=case["fmt"], hw_buff_sz=case["hw_buff_sz"]) == raw, case["stream"]
        if case["fmt"] in (q.QZ_DEFLATE_GZIP_EXT, q.QZ_DEFLATE_RAW) and len(raw) > 1:
            sess = prod.new_session(fmt=case["fmt"], hw_buff_sz=case["hw_buff_sz"])
            assert stream_decompress(prod, sess, blob, 100, 1 << 16) == raw, case["stream"]
            prod.end_session(sess)


def test_stream_decompress_raw_piecemeal(prod, ref, data):
    """reference test mode 9 on QZ_DEFLATE_RAW (test/main.c:2506-2848; piecemeal path src/qatzip_stream.c:599-749): a raw
    stream of several chunks comes back through qzDecompressStream in slices -- hw_buff_sz / 4 as the reference feeds it, and a
    256-byte drip -- without the whole stream ever being staged at once being a requirement: everything up to the last flush
    marker that has arrived is delivered"""
    d = pick(data, 700000, 21)
    # made by our qzCompressStream, and by the reference's software path in one call
    sess = prod.new_session(fmt=q.QZ_DEFLATE_RAW)
    ours = stream_compress(prod, sess, d, 30000, 1 << 20)[0]
    prod.end_session(sess)
    theirs = ref.compress(d, fmt=q.QZ_DEFLATE_RAW)
    assert zlib.decompress(ours, -15) == d
    for blob in (ours, theirs):
        for slice_sz, out_sz in ((16384, 1 << 20), (256, 1 << 20), (100000, 65536)):
            sess = prod.new_session(fmt=q.QZ_DEFLATE_RAW)
            assert stream_decompress(prod, sess, blob, slice_sz, out_sz) == d
            prod.end_session(sess)


def test_stream_rejects_other_formats(prod):
    sess = prod.new_session(fmt=q.QZ_DEFLATE_GZIP)
    st = q.QzStream(); buf = (C.c_ubyte * 4096)()
    st.in_ = C.addressof(buf); st.in_sz = 4096; st.out = C.addressof(buf); st.out_sz = 4096
    assert prod.lib.qzCompressStream(C.byref(sess), C.byref(st), 1) == q.QZ_PARAMS     # reference src/qatzip_stream.c:478-484
    prod.end_session(sess)


# ---------------------------------------------------------------------------- member discovery robustness
def test_gzip_magic_inside_payload(prod, ref):
    """Plain gzip members carry no size: like the reference (src/qatzip_gzip.c:244-261) the next member is
    found by scanning for 1f 8b 08 00.  Payloads that contain those bytes (stored blocks of random
    data) make the scan guess wrong; the engine must notice (length/CRC/ISIZE) and recover."""
    import random
    rnd = random.Random(7)
    chunk = bytearray(rnd.getrandbits(8) for _ in range(150000))
    for pos in (100, 5000, 70000, 70100, 140000):
        chunk[pos:pos + 10] = bytes([0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 0xff])
        if pos in (5000, 70100):
            chunk[pos - 4:pos] = (1000).to_bytes(4, "little")    # a believable ISIZE in front: passes the scan's plausibility filter
    d = bytes(chunk)
    for maker in (prod, ref):
        blob = maker.compress(d, fmt=q.QZ_DEFLATE_GZIP)
        assert blob.count(bytes([0x1f, 0x8b, 8, 0])) > 3
        assert prod.decompress(blob, len(d) + 8, fmt=q.QZ_DEFLATE_GZIP) == d
    two = prod.compress(d, fmt=q.QZ_DEFLATE_GZIP) + ref.compress(d[:70000], fmt=q.QZ_DEFLATE_GZIP)
    assert prod.decompress(two, len(d) + 70000 + 8, fmt=q.QZ_DEFLATE_GZIP) == d + d[:70000]


def test_mixed_member_sizes_like_config3(prod, ref, corpus):
    """BASELINE configs[2] in small: gzip members made by the reference software path with uncompressed
    sizes drawn from {4..256} KiB, decoded in one call with hw_buff_sz = 256 KiB."""
    data = corpus.make(q.Corpus.SILESIA_LIKE, 8 << 20, first_seg=40)
    sizes, state, pos, blob, want = [4, 8, 16, 32, 64, 128, 256], 3, 0, b"", b""
    while pos + (256 << 10) <= len(data):
        state = (state * 6364136223846793005 + 1442695040888963407) & ((1 << 64) - 1)
        n = sizes[(state >> 33) % 7] << 10
        piece = data[pos:pos + n]
        blob += ref.compress(piece, fmt=q.QZ_DEFLATE_GZIP, hw_buff_sz=262144)
        want += piece
        pos += n
    assert prod.decompress(blob, len(want) + 8, fmt=q.QZ_DEFLATE_GZIP, hw_buff_sz=262144) == want


# ---------------------------------------------------------------------------- threads and the async entry points
def test_concurrent_sessions(prod, port, data):
    """API contract: thread-safe across sessions (reference run_perf_test.sh runs one session per thread)."""
    import threading
    errs = []

    def work(t):
        try:
            fmt = ALL_FMTS[t % len(ALL_FMTS)]
            d = pick(data, 700000 + 4099 * t, t)
            for _ in range(3):
                blob = prod.compress(d, fmt=fmt)
                assert port.decompress(blob, fmt, len(d) + 8) == d
                assert prod.decompress(blob, len(d) + 8, fmt=fmt) == d
        except Exception as e:      # noqa: BLE001
            errs.append((t, repr(e)))
    ths = [threading.Thread(target=work, args=(t,)) for t in range(8)]
    [t.start() for t in ths]; [t.join() for t in ths]
    assert not errs, errs


class QzResult(C.Structure):   # reference include/qatzip.h QzResult_T
    _fields_ = [("status", C.c_int), ("cb_tag", C.c_void_p), ("src_len", C.c_uint), ("dest_len", C.c_uint),
                ("ext_rc", C.c_uint64), ("crc", C.c_void_p), ("extension_result", C.c_void_p)]


def test_async_callbacks(prod, port, data):
    """qzCompress2 / qzDecompress2 with a callback (reference test modes 28/29): requests complete in order,
    the callback sees status and the consumed / produced lengths; NULL callback = synchronous call."""
    import threading
    CB = C.CFUNCTYPE(C.c_int, C.POINTER(QzResult))
    L = prod.lib
    L.qzCompress2.argtypes = [C.POINTER(q.QzSession), C.c_void_p, C.c_void_p, CB, C.POINTER(QzResult)]
    L.qzDecompress2.argtypes = [C.POINTER(q.QzSession), C.c_void_p, C.c_void_p, CB, C.POINTER(QzResult)]
    sess = prod.new_session(fmt=q.QZ_DEFLATE_GZIP_EXT)
    done, order = threading.Semaphore(0), []

    def on_done(res):
        order.append(res.contents.cb_tag); done.release(); return 0
    cb = CB(on_done)
    n, reqs = 400000, []
    for i in range(4):
        d = pick(data, n, 31 * i)
        src, dst, r = C.create_string_buffer(d, n), C.create_string_buffer(n + 65536), QzResult()
        r.src_len, r.dest_len, r.cb_tag = n, n + 65536, i + 1
        assert L.qzCompress2(C.byref(sess), src, dst, cb, C.byref(r)) == q.QZ_OK
        reqs.append((d, src, dst, r))
    for _ in reqs:
        assert done.acquire(timeout=60)
    assert order == [1, 2, 3, 4]
    for d, src, dst, r in reqs:
        assert r.status == q.QZ_OK and r.src_len == n
        assert port.decompress(dst.raw[:r.dest_len], q.QZ_DEFLATE_GZIP_EXT, n + 8) == d
    # async decompress of the first result, then the synchronous form (callback NULL)
    d, src, dst, r = reqs[0]
    back, r2 = C.create_string_buffer(n + 8), QzResult()
    r2.src_len, r2.dest_len, r2.cb_tag = r.dest_len, n + 8, 9
    assert L.qzDecompress2(C.byref(sess), dst, back, cb, C.byref(r2)) == q.QZ_OK
    assert done.acquire(timeout=60) and r2.status == q.QZ_OK and back.raw[:r2.dest_len] == d
    r3 = QzResult(); r3.src_len, r3.dest_len = n, n + 65536
    assert L.qzCompress2(C.byref(sess), src, dst, CB(), C.byref(r3)) == q.QZ_OK and r3.status == q.QZ_OK and r3.src_len == n
    assert L.qzCompress2(C.byref(sess), src, dst, cb, None) == q.QZ_PARAMS
    prod.end_session(sess)          # drains and stops the completion thread


def test_stream_config5_pattern_driven_from_c(prod, port, data):
    """BASELINE configs[4] in small: DEFLATE_RAW, 4 KiB submissions, `last` on the final one, the loop written in C
    (harness/stream_drive.c) exactly as the perf tool times it; the bytes must be one valid raw deflate stream."""
    d = pick(data, 16 << 20, 15)
    drv = C.CDLL(os.path.join(os.path.dirname(q.CORPUS_SO), "libqzdrive.so"))
    drv.qzdrive_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint, C.c_void_p, C.c_uint, C.c_void_p, C.c_size_t,
                                   C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint), C.POINTER(C.c_double)]
    sess = prod.new_session(fmt=q.QZ_DEFLATE_RAW)
    ocap = 8 << 20
    obuf, sink = (C.c_ubyte * ocap)(), (C.c_ubyte * len(d))()
    outb, calls, crc, secs = C.c_uint64(0), C.c_uint64(0), C.c_uint(0), C.c_double(0)
    rc = drv.qzdrive_stream(C.cast(prod.lib.qzCompressStream, C.c_void_p), C.cast(prod.lib.qzEndStream, C.c_void_p), C.byref(sess), d, len(d), 4096,
                            obuf, ocap, sink, len(d), C.byref(outb), C.byref(calls), C.byref(crc), C.byref(secs))
    prod.end_session(sess)
    assert rc == q.QZ_OK and calls.value == len(d) // 4096 and crc.value == zlib.crc32(d)
    z = zlib.decompressobj(-15)
    assert z.decompress(bytes(sink[:outb.value])) == d and z.eof
