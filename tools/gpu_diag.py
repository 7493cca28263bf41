"""First-contact diagnostics on the GPU box: exercises the C ABI step by step and prints what
breaks where (tests/ has the real assertions; this prints instead of stopping)."""
import os, sys, time, zlib, traceback, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from harness import qzapi as q

os.makedirs("gpurun_out", exist_ok=True)
prod = q.QzLib(q.PRODUCT_SO); port = q.OraclePort(); cor = q.Corpus()
print("devices", prod.lib.qzb200DeviceCount(), flush=True)
fmts = [int(x) for x in os.environ.get("DIAG_FMTS", "2,1,0,3").split(",")]
data_all = cor.make(q.Corpus.SILESIA_LIKE, 12 << 20)

def check_members(blob, data, fmt, hw):
    if fmt != q.QZ_DEFLATE_GZIP_EXT: return -1, '', b''
    """walk gzip-ext members with python zlib; return index of first bad chunk or -1"""
    off = 0; i = 0; pos = 0
    while off < len(blob):
        if fmt == q.QZ_DEFLATE_GZIP_EXT:
            src_sz = int.from_bytes(blob[off+16:off+20], "little"); dst_sz = int.from_bytes(blob[off+20:off+24], "little")
            payload = blob[off+24:off+24+dst_sz]
            try:
                out = zlib.decompress(payload, -15)
            except Exception as e:
                return i, f"zlib error {e} (member at {off}, src_sz {src_sz} dst_sz {dst_sz})", payload
            if out != data[pos:pos+src_sz]:
                bad = next((k for k in range(min(len(out), src_sz)) if out[k] != data[pos+k]), -1)
                return i, f"mismatch at byte {bad} of chunk (got len {len(out)} want {src_sz})", payload
            crc = int.from_bytes(blob[off+24+dst_sz:off+28+dst_sz], "little")
            if crc != zlib.crc32(out): return i, f"crc footer {crc:08x} != {zlib.crc32(out):08x}", payload
            off += 24 + dst_sz + 8; pos += src_sz; i += 1
        else:
            return -1, "", b""
    return -1, "", b""

for fmt in fmts:
    for n in ((0, 100, 8193, 65537, 300000) if os.environ.get('DIAG_QUICK') else (0, 1, 100, 4096, 8191, 8192, 8193, 65536, 65537, 1 << 20, 12 << 20)):
        for seg in ((0, 8) if os.environ.get('DIAG_QUICK') else (0, 2, 8, 11)):
            if n > (1 << 20) and seg: continue
            data = data_all[seg << 20:(seg << 20) + n] if n <= (1 << 20) else data_all[:n]
            tag = f"fmt={q.FMT_NAMES[fmt]} n={n} seg={seg}"
            try:
                t = time.time()
                sess = prod.new_session(fmt=fmt)
                cap = n + n // 4 + 65536
                dst = bytearray(cap)
                rc, used, made, crc = prod.compress_call(sess, data, n, dst, cap, crc=0)
                dt = time.time() - t
                st = q.QzLib  # noqa
                if rc != 0 or used != n:
                    print("FAIL", tag, "rc", rc, "used", used, "made", made, flush=True); prod.end_session(sess); continue
                blob = bytes(dst[:made])
                if fmt != q.FMT_LZ4 and n and crc != zlib.crc32(data):
                    print("FAIL", tag, f"crc {crc:08x} want {zlib.crc32(data):08x}", flush=True)
                bad, why, payload = check_members(blob, data, fmt, 65536)
                if bad >= 0:
                    print("FAIL", tag, "chunk", bad, why, flush=True)
                    open(f"gpurun_out/bad_{q.FMT_NAMES[fmt]}_{n}_{seg}.bin", "wb").write(payload)
                    prod.end_session(sess); continue
                rcd, usedd, out = port.decompress_call(blob, fmt, n + 16)
                ok1 = (rcd == 0 and out == data)
                # our decompress of our stream
                d2 = bytearray(n + 16)
                rc2, used2, made2 = prod.decompress_call(sess, blob, len(blob), d2, n + 16) if n else (0, 0, 0)
                ok2 = (rc2 == 0 and bytes(d2[:made2]) == data) if n else True
                # our decompress of the oracle's stream
                theirs = port.compress(data, fmt)
                d3 = bytearray(n + 16)
                rc3, used3, made3 = prod.decompress_call(sess, theirs, len(theirs), d3, n + 16) if n else (0, 0, 0)
                ok3 = (rc3 == 0 and bytes(d3[:made3]) == data) if n else True
                prod.end_session(sess)
                print("ok  " if (ok1 and ok2 and ok3) else "FAIL", tag, f"ratio {made / max(n,1):.4f} oracle {len(theirs) / max(n,1):.4f} t={dt*1e3:.1f}ms",
                      "" if ok1 else f"[oracle-decode rc={rcd}]", "" if ok2 else f"[self-decode rc={rc2} used={used2}/{len(blob)} made={made2}]",
                      "" if ok3 else f"[decode-oracle rc={rc3} used={used3}/{len(theirs)} made={made3}]", flush=True)
            except Exception:
                print("EXC ", tag); traceback.print_exc(); sys.stdout.flush()
print("diag done", flush=True)
