"""Fuzz the kernels' logic on the SIMT emulator: random structured inputs through the deflate kernels (per piece, window) and LZ4, decoded by the oracle port / zlib and by our own inflate kernel.  Test infrastructure; run
by hand:  python tools/emu_fuzz.py [seconds] [seed]"""
import os, random, struct, sys, time, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from harness import qzemu as E
from harness.qzapi import OraclePort

E.build()
emu, port = E.Emu(), OraclePort()
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
WORDS = [bytes(rng.choice(b"abcdefghijklmnopqrstuvwxyz ,.") for _ in range(rng.randint(2, 11))) for _ in range(300)]


def gen(n):
    out = bytearray()
    while len(out) < n:
        kind = rng.random()
        if kind < 0.25:
            out += b"".join(rng.choice(WORDS) for _ in range(rng.randint(5, 400)))
        elif kind < 0.40:
            out += bytes([rng.randrange(256)]) * rng.randint(1, 5000)
        elif kind < 0.55:
            out += rng.randbytes(rng.randint(1, 9000))
        elif kind < 0.70 and out:
            d = rng.randint(1, min(len(out), 40000)); ln = rng.randint(3, 3000)
            for _ in range(ln):
                out.append(out[-d])
        elif kind < 0.85:
            rec = struct.pack("<IiiI", rng.randrange(1 << 20), rng.randint(-3, 3), 0, 0x3f800000 | rng.randrange(1 << 23))
            out += rec * rng.randint(1, 50) + b"".join(struct.pack("<I", i) + rec[4:] for i in range(rng.randint(1, 200)))
        else:
            period = rng.randbytes(rng.randint(1, 40)); out += period * rng.randint(1, 600)
    return bytes(out[:n])


def decode_any(blob, fmt, n):
    if fmt != E.FMT_ZLIB:
        return port.decompress(blob, fmt, n + 16)
    p, got = 0, b""
    while p < len(blob):
        d = zlib.decompressobj(15); got += d.decompress(blob[p:]); assert d.eof; p = len(blob) - len(d.unused_data)
    return got


t0, cases = time.time(), 0
while time.time() - t0 < budget:
    n = rng.choice([rng.randint(0, 300), rng.randint(0, 20000), rng.randint(60000, 70000), rng.randint(0, 300000)])
    data = gen(n)
    fmt = rng.choice([E.FMT_4B, E.FMT_GZIP, E.FMT_GZIP_EXT, E.FMT_RAW, E.FMT_ZLIB])
    chunk = rng.choice([65536, 65536, 131072, 524288])
    static = rng.random() < 0.1
    what = rng.random()
    tag = None
    try:
        if what < 0.55 and n:
            tag = "window"
            blob, cks = emu.deflate(data, fmt, chunk=chunk, static=int(static), window=1, hb=rng.choice([9, 10, 700, 2584]), grid=rng.randint(1, 3))
        elif what < 0.85:
            tag = "piece"
            chunk = rng.choice([1024, 4096, 16384, 65536, 131072])
            blob, cks = emu.deflate(data, fmt, chunk=chunk, static=int(static), piece_log2=rng.choice([13, 13, 14]), hb=12, warps=rng.randint(1, 6), nbuf=1 + rng.randint(0, 0), grid=rng.randint(1, 3))
        else:
            tag = "lz4"; fmt = E.FMT_LZ4; chunk = 65536
            blob, cks = emu.lz4(data, chunk=65536, warps=rng.randint(1, 4), grid=rng.randint(1, 2))
        assert decode_any(blob, fmt, n) == data, "round trip"
        want = port.xxh32 if fmt == E.FMT_LZ4 else zlib.adler32 if fmt == E.FMT_ZLIB else zlib.crc32
        assert cks == [want(data[i:i + chunk]) for i in range(0, max(n, 1), chunk)], "checksums"
        if fmt == E.FMT_GZIP_EXT and n:
            # our inflate kernel restores what our deflate kernels made
            p, members, doff = 0, [], 0
            while p < len(blob):
                src_sz, dst_sz = struct.unpack_from("<II", blob, p + 16)
                crc, isz = struct.unpack_from("<II", blob, p + 24 + dst_sz)
                members.append(dict(src_off=p + 24, src_len=dst_sz, exact_len=1, dst_off=doff, dst_cap=isz, exact_out=1, expect_cksum=crc, check_cksum=1))
                doff += isz; p += 32 + dst_sz
            out, res = emu.decode(E.FMT_GZIP_EXT, blob, members, n)
            assert all(r.status == E.ST_OK for r in res) and out == data, "our inflate"
    except AssertionError as e:
        open("/tmp/emu_fuzz_fail.bin", "wb").write(data)
        print("FAIL", tag, e, dict(n=n, fmt=fmt, chunk=chunk, static=static), "input saved to /tmp/emu_fuzz_fail.bin")
        sys.exit(1)
    cases += 1
print(f"{cases} cases ok in {time.time() - t0:.0f} s")

# ---- inflate / LZ4 decode: streams made by zlib and by the oracle, intact and with flipped bits (must never write out of
# bounds -- the process would die -- and must never report success with wrong bytes)
t1, dcases, rejected = time.time(), 0, 0
while time.time() - t1 < budget / 2:
    n = rng.choice([rng.randint(0, 300), rng.randint(0, 20000), rng.randint(0, 150000)])
    data = gen(n)
    if rng.random() < 0.8:
        c = zlib.compressobj(rng.choice([0, 1, 1, 3, 6, 9]), zlib.DEFLATED, -rng.choice([9, 12, 15]), rng.choice([1, 8, 9]),
                             rng.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED]))
        parts = []
        step = rng.choice([n + 1, 1000, 30000])
        for i in range(0, max(n, 1), step):
            parts.append(c.compress(data[i:i + step]))
            if rng.random() < 0.3:
                parts.append(c.flush(rng.choice([zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH])))
        raw = b"".join(parts) + c.flush()
        m = dict(src_off=0, src_len=len(raw), exact_len=1, dst_off=0, dst_cap=n, exact_out=1, expect_cksum=zlib.crc32(data), check_cksum=1)
        out, res = emu.decode(E.FMT_GZIP, raw, [m], n)
        assert res[0].status == E.ST_OK and out == data and res[0].consumed == len(raw), ("inflate", res[0].status)
        if raw and rng.random() < 0.7:
            bad = bytearray(raw)
            for _ in range(rng.randint(1, 3)):
                bad[rng.randrange(len(bad))] ^= 1 << rng.randrange(8)
            out, res = emu.decode(E.FMT_GZIP, bytes(bad), [m], n)
            if res[0].status == E.ST_OK:
                assert out == data, "corrupt stream accepted with wrong bytes"
            else:
                rejected += 1
    else:
        blob = port.compress(data, E.FMT_LZ4)
        p, members, doff = 0, [], 0
        while p < len(blob):
            (cs,) = struct.unpack_from("<Q", blob, p + 6); q = p + 15
            while True:
                (bs,) = struct.unpack_from("<I", blob, q)
                if bs == 0: break
                q += 4 + (bs & 0x7fffffff)
            (xxh,) = struct.unpack_from("<I", blob, q + 4)
            members.append(dict(src_off=p + 15, src_len=q - p - 15, exact_len=1, dst_off=doff, dst_cap=cs, exact_out=1, expect_cksum=xxh, check_cksum=1))
            doff += cs; p = q + 8
        out, res = emu.decode(E.FMT_LZ4, blob, members, n)
        assert all(r.status == E.ST_OK for r in res) and out == data, "lz4 decode"
        if len(blob) > 30 and rng.random() < 0.7:
            bad = bytearray(blob); lo = members[0]["src_off"]
            bad[rng.randrange(lo, len(bad) - 8)] ^= 1 << rng.randrange(8)
            out, res = emu.decode(E.FMT_LZ4, bytes(bad), members, n)
            if all(r.status == E.ST_OK for r in res):
                assert out == data, "corrupt lz4 accepted with wrong bytes"
            else:
                rejected += 1
    dcases += 1
print(f"{dcases} decode cases ok ({rejected} corrupted streams rejected) in {time.time() - t1:.0f} s")
