"""CPU checks of the CUDA kernels' warp logic through the SIMT emulator (tests/emu, harness/qzemu.py).

The kernel sources of qatzip_b200/csrc/*.cu are compiled by g++ against tests/emu/warp_emu.h and run here
one CTA at a time: what these tests exercise is the very code the GPU runs (match selection, Huffman
construction, bit emission, framing, the inflate and LZ4 decoders), checked against the oracle port and
Python's zlib.  This is test infrastructure: the product has no CPU path, nothing here is timed, and the
parity tests proper are the `-m gpu` ones that go through the C ABI on a B200.
"""
import struct
import os
import zlib

import pytest

from harness import qzemu as E
from harness.qzapi import Corpus


@pytest.fixture(scope="module")
def emu(built):
    E.build()
    return E.Emu()


def rle(n, seed=1):
    return Corpus().make(Corpus.REF_RLE, n, seed=seed)


def sil(n):
    return Corpus().make(Corpus.SILESIA_LIKE, max(n, 1))[:n]


def noise(n, seed=7):
    x, out = seed or 1, bytearray()
    while len(out) < n:
        x = (x * 6364136223846793005 + 1442695040888963407) & (2 ** 64 - 1)
        out += struct.pack("<Q", x)
    return bytes(out[:n])


def walk_gzip(blob, ext):
    """-> [(payload_off, payload_len, crc, isize, hdr_src, hdr_dst)] for a run of members made by the framing kernel"""
    out, p = [], 0
    while p < len(blob):
        assert blob[p:p + 3] == b"\x1f\x8b\x08"
        if ext:
            assert blob[p + 3] == 4 and blob[p + 10:p + 16] == b"\x0c\x00QZ\x08\x00" and blob[p + 9] == 0xff
            src_sz, dst_sz = struct.unpack_from("<II", blob, p + 16)
            h = 24
        else:
            assert blob[p + 3] == 0 and blob[p + 9] == 0xff
            # plain gzip carries no size: find the end by inflating
            d = zlib.decompressobj(-15)
            d.decompress(blob[p + 10:])
            dst_sz, src_sz, h = len(blob) - p - 10 - len(d.unused_data), None, 10
        crc, isize = struct.unpack_from("<II", blob, p + h + dst_sz)
        out.append((p + h, dst_sz, crc, isize, src_sz))
        p += h + dst_sz + 8
    assert p == len(blob)
    return out


def inflate_raw(b):
    d = zlib.decompressobj(-15)
    out = d.decompress(b) + d.flush()
    return out, d.eof, len(b) - len(d.unused_data)


SIZES = [0, 1, 31, 100, 4095, 8191, 8192, 8193, 20000, 65535, 65536, 65537, 150001]


@pytest.mark.parametrize("n", SIZES)
def test_deflate_gzip_ext_sizes(emu, port, n):
    data = sil(n)
    blob, cks = emu.deflate(data, E.FMT_GZIP_EXT)
    members = walk_gzip(blob, ext=True)
    assert len(members) == max(1, (n + 65535) // 65536)
    got = b""
    for i, (off, ln, crc, isize, src_sz) in enumerate(members):
        piece, eof, used = inflate_raw(blob[off:off + ln])
        chunk = data[i * 65536:(i + 1) * 65536]
        assert eof and used == ln and piece == chunk
        assert crc == zlib.crc32(chunk) == cks[i] and isize == len(chunk) == src_sz
        got += piece
    assert got == data
    if n == 0:
        assert len(blob) == 34          # QZ_COMPRESSED_SZ_OF_EMPTY_FILE
    # and the oracle's member walk + inflate restores it too
    assert port.decompress(blob, E.FMT_GZIP_EXT, n + 16) == data


@pytest.mark.parametrize("fmt", [E.FMT_4B, E.FMT_GZIP, E.FMT_GZIP_EXT, E.FMT_RAW, E.FMT_ZLIB])
@pytest.mark.parametrize("make", [sil, rle, noise, lambda n: b"\0" * n, lambda n: (b"abcabcabd" * (n // 9 + 1))[:n]])
def test_deflate_formats_and_corpora(emu, port, fmt, make):
    data = make(70000)
    blob, cks = emu.deflate(data, fmt, chunk=16384)
    if fmt == E.FMT_ZLIB:
        p, got = 0, b""
        for i in range(len(cks)):
            d = zlib.decompressobj(15)
            got += d.decompress(blob[p:])
            assert d.eof
            p = len(blob) - len(d.unused_data)
            assert cks[i] == zlib.adler32(data[i * 16384:(i + 1) * 16384])
        assert p == len(blob) and got == data
    else:
        assert port.decompress(blob, fmt, len(data) + 16) == data
        assert cks == [zlib.crc32(data[i:i + 16384]) for i in range(0, len(data), 16384)]
    if fmt == E.FMT_RAW:        # one continuous raw stream: non-final chunks end on a flush, the last one carries BFINAL
        out, eof, used = inflate_raw(blob)
        assert out == data and eof and used == len(blob)
    if fmt == E.FMT_4B:
        p = 0
        for i in range(len(cks)):
            (ln,) = struct.unpack_from("<I", blob, p)
            out, eof, used = inflate_raw(blob[p + 4:p + 4 + ln])
            assert eof and used == ln and out == data[i * 16384:(i + 1) * 16384]
            p += 4 + ln
        assert p == len(blob)


def test_deflate_raw_not_last_leaves_stream_open(emu):
    data = sil(40000)
    blob, _ = emu.deflate(data, E.FMT_RAW, chunk=16384, last=0)
    out, eof, used = inflate_raw(blob)
    assert out == data and not eof and used == len(blob)
    assert blob[-4:] == b"\x00\x00\xff\xff"            # ends like Z_FULL_FLUSH


@pytest.mark.parametrize("geom", [dict(piece_log2=13, hb=11, warps=20, nbuf=17, grid=1), dict(piece_log2=13, hb=11, warps=3, nbuf=1, grid=3),
                                  dict(piece_log2=13, hb=12, warps=4, nbuf=2, grid=2), dict(piece_log2=14, hb=12, warps=4, nbuf=3, grid=2),
                                  dict(piece_log2=14, hb=13, warps=2, nbuf=2, grid=1), dict(piece_log2=13, hb=11, warps=10, nbuf=8, grid=2)])
def test_deflate_geometries(emu, port, geom):
    data = sil(300000)
    blob, cks = emu.deflate(data, E.FMT_GZIP, chunk=65536, **geom)
    assert port.decompress(blob, E.FMT_GZIP, len(data) + 16) == data
    assert cks == [zlib.crc32(data[i:i + 65536]) for i in range(0, len(data), 65536)]


def test_deflate_output_does_not_depend_on_geometry(emu, port):
    """Pieces are independent: which warp, which piece buffer and how many CTAs must not change a byte.  (The kernel has
    one intended write/write race, equal-hash insertions inside one 32-position tile; the emulator always lets the highest
    lane win it, the GPU may not, which can move a few bytes per piece -- DESIGN.md section 7.)"""
    data = sil(200000)
    a, _ = emu.deflate(data, E.FMT_GZIP_EXT, warps=4, nbuf=3, grid=2)
    b, _ = emu.deflate(data, E.FMT_GZIP_EXT, warps=7, nbuf=2, grid=1)
    c, _ = emu.deflate(data, E.FMT_GZIP_EXT, warps=20, nbuf=17, grid=3)
    assert a == b == c
    g1, _ = emu.deflate(data, E.FMT_GZIP_EXT, warps=16, nbuf=1, grid=2, window=1, hb=10)
    g2, _ = emu.deflate(data, E.FMT_GZIP_EXT, warps=32, nbuf=2, grid=1, window=1, hb=10)
    assert g1 == g2
    h1, _ = emu.deflate(data, E.FMT_GZIP_EXT, warps=16, nbuf=1, grid=2, window=1, hb=2584)
    h2, _ = emu.deflate(data, E.FMT_GZIP_EXT, warps=32, nbuf=2, grid=1, window=1, hb=2584)
    assert h1 == h2 and port.decompress(h1, E.FMT_GZIP_EXT, len(data) + 16) == data
    assert port.decompress(g1, E.FMT_GZIP_EXT, len(data) + 16) == data


@pytest.mark.parametrize("chunk", [1024, 4096, 65536, 524288])
def test_deflate_chunk_sizes(emu, port, chunk):
    data = sil(3 * chunk // 2 + 77) if chunk <= 65536 else sil(chunk + 5000)
    blob, _ = emu.deflate(data, E.FMT_GZIP_EXT, chunk=chunk)
    assert port.decompress(blob, E.FMT_GZIP_EXT, len(data) + 16) == data


def test_deflate_static_huffman_uses_fixed_or_stored_blocks(emu):
    data = sil(30000)
    blob, _ = emu.deflate(data, E.FMT_RAW, chunk=65536, static=1)
    out, eof, used = inflate_raw(blob)
    assert out == data and eof
    assert (blob[0] >> 1) & 3 in (0, 1)             # first block: stored or fixed, never dynamic


def test_deflate_ratio_near_zlib1(emu):
    data = sil(1 << 20)
    blob, _ = emu.deflate(data, E.FMT_GZIP_EXT)
    ref = sum(len(zlib.compress(data[i:i + 65536], 1)) + 20 for i in range(0, len(data), 65536))
    assert len(blob) <= 1.05 * ref, (len(blob), ref)


def test_deflate_dest_too_small_keeps_whole_chunks(emu):
    data = sil(200000)
    full, _ = emu.deflate(data, E.FMT_GZIP_EXT)
    members = walk_gzip(full, ext=True)
    two = members[2][0] - 24                      # bytes of the first two members
    part, _ = emu.deflate(data, E.FMT_GZIP_EXT, cap=two + 100)
    assert part == full[:two]


# ------------------------------------------------------------------ inflate

def one_member(payload_len, out_len, crc, src_off=0, dst_off=0, exact_len=1, exact_out=1, check=1):
    return dict(src_off=src_off, src_len=payload_len, exact_len=exact_len, dst_off=dst_off, dst_cap=out_len, exact_out=exact_out,
                expect_cksum=crc, check_cksum=check)


@pytest.mark.parametrize("level,strategy", [(0, 0), (1, 0), (6, 0), (9, 0), (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE)])
@pytest.mark.parametrize("make", [sil, rle, noise])
def test_inflate_zlib_made_streams(emu, level, strategy, make):
    data = make(100000)
    c = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
    raw = c.compress(data) + c.flush()
    out, res = emu.decode(E.FMT_GZIP, raw, [one_member(len(raw), len(data), zlib.crc32(data))], len(data))
    r = res[0]
    assert (r.status, r.consumed, r.produced, r.cksum, r.saw_final) == (E.ST_OK, len(raw), len(data), zlib.crc32(data), 1)
    assert out == data


def test_inflate_many_members_mixed_sizes(emu):
    sizes = [4096, 0, 262144, 1, 8192, 65536, 131072, 16384, 300]
    datas = [sil(sum(sizes))[sum(sizes[:i]):sum(sizes[:i + 1])] for i in range(len(sizes))]
    src, members, doff = b"", [], 0
    for d in datas:
        c = zlib.compressobj(1, zlib.DEFLATED, -15)
        raw = c.compress(d) + c.flush()
        members.append(one_member(len(raw), len(d), zlib.crc32(d), src_off=len(src), dst_off=doff))
        src += raw + b"\x55" * 3            # unrelated bytes between payloads (footers/headers in a real stream)
        doff += len(d)
    out, res = emu.decode(E.FMT_GZIP, src, members, doff, grid=3)
    assert all(r.status == E.ST_OK for r in res)
    assert out == b"".join(datas)


def test_inflate_our_own_streams(emu):
    data = sil(200000)
    blob, _ = emu.deflate(data, E.FMT_GZIP_EXT)
    members = [one_member(ln, isize, crc, src_off=off, dst_off=i * 65536) for i, (off, ln, crc, isize, _) in enumerate(walk_gzip(blob, ext=True))]
    out, res = emu.decode(E.FMT_GZIP_EXT, blob, members, len(data))
    assert [r.status for r in res] == [E.ST_OK] * len(members) and out == data


def test_inflate_zlib_format_checks_adler(emu):
    data = sil(50000)
    z = zlib.compress(data, 1)
    m = one_member(len(z) - 6, len(data), zlib.adler32(data), src_off=2)
    out, res = emu.decode(E.FMT_ZLIB, z, [m], len(data))
    assert res[0].status == E.ST_OK and out == data
    m["expect_cksum"] ^= 1
    _, res = emu.decode(E.FMT_ZLIB, z, [m], len(data))
    assert res[0].status == E.ST_CKSUM


def test_inflate_errors(emu):
    data = sil(60000)
    c = zlib.compressobj(1, zlib.DEFLATED, -15)
    raw = c.compress(data) + c.flush()
    crc = zlib.crc32(data)
    # wrong CRC, wrong ISIZE, output too small, truncated input, garbage, reserved block type
    assert emu.decode(E.FMT_GZIP, raw, [one_member(len(raw), len(data), crc ^ 1)], len(data))[1][0].status == E.ST_CKSUM
    assert emu.decode(E.FMT_GZIP, raw, [one_member(len(raw), len(data) + 1, crc)], len(data) + 1)[1][0].status == E.ST_SIZE
    assert emu.decode(E.FMT_GZIP, raw, [one_member(len(raw), len(data) - 100, crc, exact_out=0)], len(data))[1][0].status == E.ST_OUT_FULL
    assert emu.decode(E.FMT_GZIP, raw, [one_member(len(raw) // 2, len(data), crc)], len(data))[1][0].status in (E.ST_IN_TRUNC, E.ST_DATA_ERROR)
    bad = bytearray(raw); bad[len(bad) // 2] ^= 0x5a
    assert emu.decode(E.FMT_GZIP, bytes(bad), [one_member(len(raw), len(data), crc)], len(data))[1][0].status != E.ST_OK
    assert emu.decode(E.FMT_GZIP, b"\x07" + raw, [one_member(len(raw) + 1, len(data), crc)], len(data))[1][0].status == E.ST_DATA_ERROR
    # payload followed by extra bytes: exact_len catches it, without exact_len the decoder reports where it stopped
    r = emu.decode(E.FMT_GZIP, raw + b"\0" * 9, [one_member(len(raw) + 9, len(data), crc)], len(data))[1][0]
    assert r.status == E.ST_DATA_ERROR
    r = emu.decode(E.FMT_GZIP, raw + b"\0" * 9, [one_member(len(raw) + 9, len(data), crc, exact_len=0)], len(data))[1][0]
    assert r.status == E.ST_OK and r.consumed == len(raw)


def test_inflate_size_only_mode_reports_lengths_without_writing(emu):
    data = sil(50000)
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    raw = c.compress(data) + c.flush()
    out, res = emu.decode(E.FMT_ZLIB, raw + b"tail", [one_member(len(raw) + 4, 1 << 20, 0, exact_len=0, exact_out=0, check=0)], 64, size_only=1)
    assert (res[0].status, res[0].consumed, res[0].produced) == (E.ST_OK, len(raw), len(data))
    assert out == b"\0" * 64


def test_inflate_raw_chunk_without_bfinal(emu):
    data = sil(30000)
    c = zlib.compressobj(1, zlib.DEFLATED, -15)
    raw = c.compress(data) + c.flush(zlib.Z_FULL_FLUSH)
    out, res = emu.decode(E.FMT_RAW, raw, [one_member(len(raw), len(data), 0, exact_out=0, check=0)], len(data))
    assert (res[0].status, res[0].produced, res[0].saw_final) == (E.ST_OK, len(data), 0) and out == data


# ------------------------------------------------------------------ LZ4

def walk_lz4(blob):
    """-> [(payload_off, payload_len, content_size, xxh)] per frame written by the framing kernel"""
    out, p = [], 0
    while p < len(blob):
        assert blob[p:p + 6] == b"\x04\x22\x4d\x18\x4c\x40"
        (csize,) = struct.unpack_from("<Q", blob, p + 6)
        q = p + 15
        while True:
            (bs,) = struct.unpack_from("<I", blob, q)
            if bs == 0:
                break
            q += 4 + (bs & 0x7fffffff)
        (xxh,) = struct.unpack_from("<I", blob, q + 4)
        out.append((p + 15, q - p - 15, csize, xxh))
        p = q + 8
    assert p == len(blob)
    return out


@pytest.mark.parametrize("n", [0, 1, 12, 13, 100, 8192, 8193, 65536, 65537, 200000])
@pytest.mark.parametrize("make", [sil, rle, noise])
def test_lz4_round_trip(emu, port, n, make):
    data = make(n)
    blob, cks = emu.lz4(data)
    assert port.decompress(blob, E.FMT_LZ4, n + 16) == data
    frames = walk_lz4(blob)
    assert [f[2] for f in frames] == [len(data[i:i + 65536]) for i in range(0, max(n, 1), 65536)]
    assert [f[3] for f in frames] == cks == [port.xxh32(data[i:i + 65536]) for i in range(0, max(n, 1), 65536)]
    members = [one_member(ln, cs, x, src_off=off, dst_off=i * 65536) for i, (off, ln, cs, x) in enumerate(frames)]
    out, res = emu.decode(E.FMT_LZ4, blob, members, n)
    assert [r.status for r in res] == [E.ST_OK] * len(frames) and out == data


def test_lz4_decodes_oracle_frames_and_rejects_corruption(emu, port):
    data = sil(150000)
    blob = port.compress(data, E.FMT_LZ4)
    frames = walk_lz4(blob)
    members, doff = [], 0
    for off, ln, cs, x in frames:
        members.append(one_member(ln, cs, x, src_off=off, dst_off=doff)); doff += cs
    out, res = emu.decode(E.FMT_LZ4, blob, members, len(data))
    assert [r.status for r in res] == [E.ST_OK] * len(frames) and out == data
    members[0]["expect_cksum"] ^= 4
    _, res = emu.decode(E.FMT_LZ4, blob, members, len(data))
    assert res[0].status == E.ST_CKSUM and res[1].status == E.ST_OK


# ------------------------------------------------------------------ window kernel: 64 KiB windows in shared memory, one deflate block per window

def decode_any(port, blob, fmt, n):
    if fmt != E.FMT_ZLIB:
        return port.decompress(blob, fmt, n + 16)
    p, got = 0, b""
    while p < len(blob):
        d = zlib.decompressobj(15)
        got += d.decompress(blob[p:])
        assert d.eof
        p = len(blob) - len(d.unused_data)
    return got


@pytest.mark.parametrize("fmt", [E.FMT_4B, E.FMT_GZIP, E.FMT_GZIP_EXT, E.FMT_RAW, E.FMT_ZLIB])
@pytest.mark.parametrize("name,make,n", [("sil", sil, 300000), ("one_chunk", sil, 65536), ("rle", rle, 150000), ("noise", noise, 70000),
                                         ("zeros", lambda n: b"\0" * n, 100000), ("tiny", sil, 37), ("one_byte", sil, 1),
                                         ("piece_edge", sil, 65536 + 8192), ("mixed", lambda n: sil(30000) + noise(40000) + sil(n - 70000), 170000)])
@pytest.mark.parametrize("hb", [11, 2584])
def test_window_deflate_round_trip(emu, port, fmt, name, make, n, hb):
    data = make(n)
    blob, cks = emu.deflate(data, fmt, warps=16, grid=2, window=1, hb=hb)
    assert decode_any(port, blob, fmt, n) == data
    want = zlib.adler32 if fmt == E.FMT_ZLIB else zlib.crc32
    assert cks == [want(data[i:i + 65536]) for i in range(0, n, 65536)]
    if fmt == E.FMT_RAW:
        out, eof, used = inflate_raw(blob)
        assert out == data and eof and used == len(blob)


@pytest.mark.parametrize("chunk", [65536, 131072, 524288])
@pytest.mark.parametrize("geom", [dict(window=1, grid=3), dict(window=1, grid=1, hb=10), dict(window=1, grid=1, hb=11), dict(window=1, grid=2, hb=2584),
                                  dict(window=1, grid=1, hb=9), dict(window=1, grid=2, hb=700)])
def test_window_deflate_geometries_and_chunks(emu, port, chunk, geom):
    data = sil(chunk + chunk // 2 + 4321)
    blob, cks = emu.deflate(data, E.FMT_GZIP_EXT, chunk=chunk, **geom)
    assert port.decompress(blob, E.FMT_GZIP_EXT, len(data) + 16) == data
    assert cks == [zlib.crc32(data[i:i + chunk]) for i in range(0, len(data), chunk)]
    members = walk_gzip(blob, ext=True)
    assert [m[4] for m in members] == [len(data[i:i + chunk]) for i in range(0, len(data), chunk)]


def test_window_deflate_one_block_per_64k_and_smaller_than_per_piece(emu):
    data = sil(1 << 20)
    grouped, _ = emu.deflate(data, E.FMT_RAW, warps=16, nbuf=4, grid=2, window=1)
    pieces, _ = emu.deflate(data, E.FMT_RAW)
    assert inflate_raw(grouped)[0] == data
    assert len(grouped) < len(pieces)                  # 7 of 8 block headers and flush markers are gone
    ref = sum(len(zlib.compress(data[i:i + 65536], 1)) for i in range(0, len(data), 65536))
    assert len(grouped) <= 1.04 * ref, (len(grouped), ref)
    # flush markers: one per chunk boundary, not one per piece
    assert grouped.count(b"\x00\x00\xff\xff") < pieces.count(b"\x00\x00\xff\xff") // 4


def test_window_deflate_static_and_not_last(emu):
    data = sil(100000)
    blob, _ = emu.deflate(data, E.FMT_RAW, static=1, warps=16, nbuf=2, grid=1, window=1)
    out, eof, _ = inflate_raw(blob)
    assert out == data and eof and (blob[0] >> 1) & 3 in (0, 1)
    blob, _ = emu.deflate(data, E.FMT_RAW, last=0, warps=16, nbuf=2, grid=1, window=1)
    out, eof, used = inflate_raw(blob)
    assert out == data and not eof and used == len(blob) and blob[-4:] == b"\x00\x00\xff\xff"


def test_window_deflate_dest_too_small_keeps_whole_chunks(emu):
    data = sil(200000)
    full, _ = emu.deflate(data, E.FMT_GZIP_EXT, warps=16, nbuf=3, grid=2, window=1)
    members = walk_gzip(full, ext=True)
    two = members[2][0] - 24
    part, _ = emu.deflate(data, E.FMT_GZIP_EXT, cap=two + 100, warps=16, nbuf=3, grid=2, window=1)
    assert part == full[:two]


def test_window_streams_decode_with_our_inflate(emu):
    data = sil(200000)
    blob, _ = emu.deflate(data, E.FMT_GZIP_EXT, warps=16, nbuf=3, grid=2, window=1)
    members = [one_member(ln, isize, crc, src_off=off, dst_off=i * 65536) for i, (off, ln, crc, isize, _) in enumerate(walk_gzip(blob, ext=True))]
    out, res = emu.decode(E.FMT_GZIP_EXT, blob, members, len(data))
    assert [r.status for r in res] == [E.ST_OK] * len(members) and out == data


def test_window_ratio_gate(emu):
    """the ratio gate of the GPU suite (<= 1.05 x zlib -1 per 64 KiB chunk) on the emulator, at the window kernel's default
    table size: SILESIA-LIKE, REF-RLE and the real-text sample"""
    text = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "text_sample.txt"), "rb").read()
    for name, data in (("sil", sil(1 << 20)), ("rle", rle(1 << 19)), ("text", text)):
        ours, _ = emu.deflate(data, E.FMT_RAW, warps=32, grid=2, window=1, hb=2584)
        ref = sum(len(zlib.compress(data[i:i + 65536], 1)) - 6 for i in range(0, len(data), 65536))
        assert inflate_raw(ours)[0] == data
        assert len(ours) <= 1.05 * ref, (name, len(ours), ref)


# ------------------------------------------------------------------ LZ4 window kernel: one 64 KiB block per chunk

@pytest.mark.parametrize("nw,tent", [(12, 6900), (16, 5200), (12, 300)])
@pytest.mark.parametrize("name,make,n", [("sil", sil, 300000), ("one_chunk", sil, 65536), ("rle", rle, 150000), ("noise", noise, 70000), ("zeros", lambda n: b"\0" * n, 200000),
                                         ("tiny", sil, 11), ("twelve", sil, 12), ("thirteen", lambda n: b"a" * n, 13), ("edge", sil, 65536 + 4097)])
def test_lz4_window_round_trip(emu, port, name, make, n, nw, tent):
    data = make(n)
    blob, cks = emu.lz4_window(data, tent=tent, nw=nw)
    assert port.decompress(blob, E.FMT_LZ4, n + 16) == data
    assert cks == [port.xxh32(data[i:i + 65536]) for i in range(0, n, 65536)]
    frames = walk_lz4(blob)
    assert len(frames) == (n + 65535) // 65536                 # one frame per chunk ...
    members = [dict(src_off=off, src_len=ln, exact_len=1, dst_off=i * 65536, dst_cap=size, exact_out=1, expect_cksum=ck, check_cksum=1) for i, (off, ln, size, ck) in enumerate(frames)]
    out, res = emu.decode(E.FMT_LZ4, blob, members, n)          # ... that our own decoder takes back
    assert [r.status for r in res] == [E.ST_OK] * len(frames) and out == data


def test_lz4_window_larger_chunks_and_ratio(emu, port):
    data = sil(400000)
    blob, _ = emu.lz4_window(data, chunk=131072)
    assert port.decompress(blob, E.FMT_LZ4, len(data) + 16) == data
    text = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "text_sample.txt"), "rb").read()
    for name, d in (("sil", sil(1 << 20)), ("rle", rle(1 << 19)), ("text", text)):
        ours, _ = emu.lz4_window(d)
        ref = sum(len(port.compress(d[i:i + 65536], E.FMT_LZ4)) for i in range(0, len(d), 65536))
        assert len(ours) <= 1.05 * ref, (name, len(ours), ref)


def test_window_deeper_search_for_higher_levels(emu, port):
    """compression levels 6 and up look at two entries per hash bucket (window=2): same decoder, smaller output on text"""
    text = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "text_sample.txt"), "rb").read()
    for data in (text, sil(300000), rle(100000), noise(70000), b"x" * 5, sil(65536 + 77)):
        one, _ = emu.deflate(data, E.FMT_GZIP_EXT, grid=2, window=1, hb=2584)
        two, cks = emu.deflate(data, E.FMT_GZIP_EXT, grid=2, window=2, hb=2584)
        assert port.decompress(two, E.FMT_GZIP_EXT, len(data) + 16) == data
        assert cks == [zlib.crc32(data[i:i + 65536]) for i in range(0, len(data), 65536)]
        if data is text:
            assert len(two) < len(one)


def test_gzip_member_scan_kernel(emu):
    """the device-side scan for member starts (device-resident decompress of plain gzip) against the host's predicate
    (qz_engine.cu gzip_member_start): magic, method, reserved flag bits clear, XFL in {0, 2, 4}, OS <= 13 or 255"""
    import gzip, random
    rng = random.Random(5)
    def plausible(p, q):
        return (q + 10 <= len(p) and p[q] == 0x1f and p[q + 1] == 0x8b and p[q + 2] == 8 and (p[q + 3] & 0xe0) == 0 and
                p[q + 8] in (0, 2, 4) and (p[q + 9] <= 13 or p[q + 9] == 255))
    members = [gzip.compress(sil(rng.randrange(1, 40000)), 1) for _ in range(9)]
    decoys = [bytes([0x1f, 0x8b, 8, 0xe0, 0, 0, 0, 0, 0, 3]), bytes([0x1f, 0x8b, 9, 0, 0, 0, 0, 0, 0, 3]), bytes([0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 1, 3]),
              bytes([0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 4, 255]), bytes([0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 14]), bytes([0x1f, 0x8b, 8])]
    blob = b"".join(m + rng.choice(decoys) for m in members) + bytes([0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0])      # a header cut short at the end
    for lo in (0, 1, len(members[0]) + 3):
        want = [q for q in range(lo, len(blob)) if plausible(blob, q)]
        got, found = emu.gzip_scan(blob, lo=lo)
        assert found == len(want) and got == want
    got, found = emu.gzip_scan(blob, cap=2)          # a list that is too short is cut, the count still says so
    assert found == len([q for q in range(len(blob)) if plausible(blob, q)]) and len(got) == 2
