"""Pins the oracle.  The C restatement (oracle/qz_oracle.c) is checked against
  * the unmodified reference compiled for the CPU (oracle/_ref), both directions, all formats;
  * golden fixtures generated from oracle/_ref (tests/golden, made by tests/golden/make_golden.py);
  * zlib crc32 / gzip(1) / known xxHash32 vectors.
The reference itself holds no golden vectors: every assertion there is a
round trip, a length, a return code or a CRC equality; those are what is restated here."""
import gzip
import hashlib
import json
import os
import zlib

import pytest

from harness import qzapi as q

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SIZES = [0, 1, 127, 1023, 1024, 4096, 65535, 65536, 65537, 200001]
FMTS = [q.QZ_DEFLATE_4B, q.QZ_DEFLATE_GZIP, q.QZ_DEFLATE_GZIP_EXT, q.QZ_DEFLATE_RAW, q.FMT_LZ4, q.FMT_ZLIB]


def sample(corpus, n, kind=q.Corpus.SILESIA_LIKE, seg=0):
    return corpus.make(kind, max(n, 1), first_seg=seg)[:n]


def test_crc32_and_combine(port, corpus):
    d = sample(corpus, 300000)
    assert port.crc32(d) == zlib.crc32(d)
    for cut in (0, 1, 1000, 65536, 299999, 300000):
        a, b = d[:cut], d[cut:]
        assert port.crc32_combine(zlib.crc32(a), zlib.crc32(b), len(b)) == zlib.crc32(d)


def test_adler32_and_zlib_streams(port, corpus):
    """zlib-format sessions: Adler-32 trailer, 78 9C header, one stream per chunk, and every stream is
    what Python's zlib (RFC 1950) reads."""
    d = sample(corpus, 300000)
    assert port.adler32(b"") == 1 and port.adler32(d) == zlib.adler32(d)
    blob = port.compress(d, q.FMT_ZLIB, hw_buff_sz=65536)
    out, rest, n = b"", blob, 0
    while rest:
        assert rest[:2] == b"\x78\x9c"
        z = zlib.decompressobj()
        out += z.decompress(rest)
        assert z.eof
        rest = z.unused_data
        n += 1
    assert out == d and n == (len(d) + 65535) // 65536


def test_xxh32_known_answers(port):
    # published xxHash32 test vectors (seed 0 and a prime seed)
    assert port.xxh32(b"") == 0x02CC5D05
    assert port.xxh32(b"", 0x9E3779B1) == 0x36B78AE7
    assert port.xxh32(b"a") == 0x550D7456
    assert port.xxh32(b"abc") == 0x32D153FF
    assert port.xxh32(b"Nobody inspects the spammish repetition") == 0xE2293B2F


@pytest.mark.parametrize("fmt", FMTS)
def test_port_vs_reference_both_directions(port, ref, corpus, fmt):
    for n in SIZES:
        d = sample(corpus, n, seg=n % 7)
        ours = port.compress(d, fmt)
        assert port.decompress(ours, fmt, n + 8) == d
        if n:   # the reference returns early for *src_len == 0 on both paths
            assert ref.decompress(ours, n + 8, fmt=fmt) == d, "reference could not decode the port's framing"
        theirs = ref.compress(d, fmt=fmt) if n else b""
        if n:
            assert port.decompress(theirs, fmt, n + 8) == d, "port could not decode the reference's stream"


def test_gzip_members_readable_by_gzip_module(port, corpus):
    d = sample(corpus, 150000, seg=3)
    for fmt in (q.QZ_DEFLATE_GZIP, q.QZ_DEFLATE_GZIP_EXT):
        assert gzip.decompress(port.compress(d, fmt)) == d


def test_framing_bytes(port, corpus):
    """Header / footer byte layouts (reference src/qatzip_gzip.c:98-143,228-237, src/qatzip_lz4.c:104-143)."""
    d = sample(corpus, 70000, seg=1)
    ext = port.compress(d, q.QZ_DEFLATE_GZIP_EXT)
    assert ext[:16] == bytes([0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 12, 0, ord("Q"), ord("Z"), 8, 0])
    src_sz, dst_sz = int.from_bytes(ext[16:20], "little"), int.from_bytes(ext[20:24], "little")
    assert src_sz == 65536
    ftr = ext[24 + dst_sz:24 + dst_sz + 8]
    assert int.from_bytes(ftr[:4], "little") == zlib.crc32(d[:65536]) and int.from_bytes(ftr[4:], "little") == 65536
    assert ext[24 + dst_sz + 8:24 + dst_sz + 12] == bytes([0x1f, 0x8b, 8, 4])
    std = port.compress(d, q.QZ_DEFLATE_GZIP)
    assert std[:10] == bytes([0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 0xff])
    b4 = port.compress(d, q.QZ_DEFLATE_4B)
    n0 = int.from_bytes(b4[:4], "little")
    assert zlib.decompress(b4[4:4 + n0], -15) == d[:65536]
    lz = port.compress(d, q.FMT_LZ4)
    assert lz[:6] == bytes([0x04, 0x22, 0x4d, 0x18, 0x4c, 0x40]) and int.from_bytes(lz[6:14], "little") == 65536
    assert lz[14] == (port.xxh32(lz[4:14]) >> 8) & 0xff
    assert len(port.compress(b"", q.QZ_DEFLATE_GZIP_EXT)) == 34      # QZ_COMPRESSED_SZ_OF_EMPTY_FILE


def test_crc_accumulates_like_hw_path(port, corpus):
    d = sample(corpus, 300000, seg=2)
    _, crc = port.compress(d, q.QZ_DEFLATE_GZIP_EXT, want_crc=True)
    assert crc == zlib.crc32(d)


def test_reference_compress_crc_single_chunk(ref, corpus):
    """reference test/main.c:4283-4337: qzCompressCrc == crc32() for 64 KiB and 1023 B."""
    for n in (65536, 1023):
        d = sample(corpus, n)
        sess = ref.new_session(fmt=q.QZ_DEFLATE_GZIP_EXT)
        dst = bytearray(n * 2 + 1024)
        rc, used, made, crc = ref.compress_call(sess, d, n, dst, len(dst), crc=0)
        ref.end_session(sess)
        assert rc == q.QZ_OK and used == n and crc == zlib.crc32(d)


def test_port_error_codes(port, corpus):
    d = sample(corpus, 100000)
    blob = bytearray(port.compress(d, q.QZ_DEFLATE_GZIP_EXT))
    bad = bytes([0x1e]) + bytes(blob[1:])
    assert port.decompress_call(bad, q.QZ_DEFLATE_GZIP_EXT, 200000)[0] == q.QZ_FAIL            # corrupt id1
    blob2 = bytearray(blob); blob2[100:140] = bytes(40)
    assert port.decompress_call(bytes(blob2), q.QZ_DEFLATE_GZIP_EXT, 200000)[0] == q.QZ_DATA_ERROR
    rc, used, out = port.decompress_call(bytes(blob), q.QZ_DEFLATE_GZIP_EXT, 1024)            # dest too small
    assert rc == q.QZ_BUF_ERROR and used == 0
    rc, used, out = port.decompress_call(bytes(blob[:-5]), q.QZ_DEFLATE_GZIP_EXT, 200000)      # truncated tail member
    assert rc == q.QZ_DATA_ERROR and out == d[:65536]


def test_golden_fixtures(port, ref):
    """Vectors generated here from the compiled reference (tests/golden/make_golden.py)."""
    with open(os.path.join(GOLD, "manifest.json")) as f:
        man = json.load(f)
    assert man["cases"], "no golden cases"
    for case in man["cases"]:
        raw = open(os.path.join(GOLD, case["input"]), "rb").read()
        blob = open(os.path.join(GOLD, case["stream"]), "rb").read()
        assert hashlib.sha256(raw).hexdigest() == case["input_sha256"]
        assert port.decompress(blob, case["fmt"], len(raw) + 8) == raw
        assert port.crc32(raw) == case["crc32"] and port.xxh32(raw) == case["xxh32"] and port.adler32(raw) == case["adler32"]
        assert ref.decompress(blob, len(raw) + 8, fmt=case["fmt"], hw_buff_sz=case["hw_buff_sz"]) == raw
