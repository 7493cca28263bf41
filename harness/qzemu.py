"""ctypes view of tests/emu/libqzemu.so -- TEST INFRASTRUCTURE.

libqzemu.so is the kernel source of qatzip_b200/csrc/*.cu compiled by g++ against a SIMT emulator
(tests/emu/warp_emu.h).  It lets the CPU suite run the kernels' warp logic without a GPU.  It is not
the product, is not loaded by the product, and is never timed.
"""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_SO = os.environ.get("QZ_EMU_SO") or os.path.join(EMU_DIR, "libqzemu.so")

# QzbFormat / QzbStatus (qatzip_b200/csrc/qz_hd.h)
FMT_4B, FMT_GZIP, FMT_GZIP_EXT, FMT_RAW, FMT_LZ4, FMT_ZLIB = 0, 1, 2, 3, 4, 5
ST_OK, ST_DATA_ERROR, ST_OUT_FULL, ST_IN_TRUNC, ST_CKSUM, ST_SIZE = 0, 1, 2, 3, 4, 5


class Member(C.Structure):  # QzbMember, qatzip_b200/csrc/qz_kernels.cuh
    _fields_ = [("src_off", C.c_uint64), ("src_len", C.c_uint32), ("exact_len", C.c_uint32), ("dst_off", C.c_uint64),
                ("dst_cap", C.c_uint32), ("exact_out", C.c_uint32), ("expect_cksum", C.c_uint32), ("check_cksum", C.c_uint32)]


class MemberResult(C.Structure):  # QzbMemberResult
    _fields_ = [("status", C.c_uint32), ("consumed", C.c_uint32), ("produced", C.c_uint32), ("cksum", C.c_uint32),
                ("saw_final", C.c_uint32), ("safe_consumed", C.c_uint32), ("safe_produced", C.c_uint32), ("pad", C.c_uint32)]


def build():
    subprocess.run(["make", "-C", EMU_DIR, "-j4"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.STDOUT)


class Emu:
    def __init__(self, path=EMU_SO):
        self.lib = L = C.CDLL(path)
        V = C.c_void_p
        L.emu_deflate_compress.restype = C.c_long
        L.emu_deflate_compress.argtypes = [C.c_int, V, C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                           V, C.c_uint64, V, C.c_int]
        L.emu_lz4_compress.restype = C.c_long
        L.emu_lz4_compress.argtypes = [V, C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_int, V, C.c_uint64, V]
        L.emu_lz4_window.restype = C.c_long
        L.emu_lz4_window.argtypes = [V, C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_int, V, C.c_uint64, V]
        L.emu_inflate.argtypes = [C.c_int, V, V, C.POINTER(Member), C.POINTER(MemberResult), C.c_uint32, C.c_int, C.c_int]
        L.emu_gzip_scan.argtypes = [V, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint32), C.c_uint32, C.c_int]
        L.emu_lz4_decompress.argtypes = [V, V, C.POINTER(Member), C.POINTER(MemberResult), C.c_uint32, C.c_int]

    def deflate(self, data, fmt=FMT_GZIP_EXT, chunk=65536, last=1, static=0, piece_log2=13, hb=11, warps=4, nbuf=3, grid=2, cap=None, window=0):
        """-> (stream bytes, [per-chunk checksum]).  window=8 / 16: the window kernel with that many warps per 64 KiB window (the CTA's groups share nbuf units; hb >= 256 is the
        number of table entries itself, else its log2)"""
        data = bytes(data)
        nch = max(1, (len(data) + chunk - 1) // chunk)
        cap = cap if cap is not None else len(data) + len(data) // 8 + 512 * nch + 64
        dst = C.create_string_buffer(max(cap, 1) + 4096)
        ck = (C.c_uint32 * nch)()
        n = self.lib.emu_deflate_compress(fmt, data, len(data), chunk, last, static, piece_log2, hb, warps, nbuf, grid, dst, cap, ck, window)
        assert n != -2, "a kernel wrote past its scratch buffers"
        assert n >= 0, "geometry not offered by the kernel"
        assert dst.raw[max(cap, 1):] == b"\0" * 4096, "the framing kernel wrote past the destination"
        return dst.raw[:n], list(ck)

    def lz4(self, data, chunk=65536, piece_log2=13, warps=4, grid=2):
        data = bytes(data)
        nch = max(1, (len(data) + chunk - 1) // chunk)
        cap = len(data) + len(data) // 8 + 512 * nch + 64
        dst = C.create_string_buffer(cap)
        ck = (C.c_uint32 * nch)()
        n = self.lib.emu_lz4_compress(data, len(data), chunk, piece_log2, warps, grid, dst, cap, ck)
        assert n >= 0
        return dst.raw[:n], list(ck)

    def lz4_window(self, data, chunk=65536, tent=6900, nw=12, grid=2):
        """window kernel: one LZ4 block per 64 KiB -> (frames, [XXH32 per chunk])"""
        data = bytes(data)
        nch = max(1, (len(data) + chunk - 1) // chunk)
        cap = len(data) + len(data) // 8 + 512 * nch + 64
        dst = C.create_string_buffer(cap)
        ck = (C.c_uint32 * nch)()
        n = self.lib.emu_lz4_window(data, len(data), chunk, tent, nw, grid, dst, cap, ck)
        assert n >= 0
        return dst.raw[:n], list(ck)

    def gzip_scan(self, src, lo=0, cap=4096, grid=3):
        """offsets (absolute, sorted) of plausible gzip member starts in src[lo:], and the number found"""
        src = bytes(src)
        lst = (C.c_uint32 * cap)()
        found = self.lib.emu_gzip_scan(src + b"\0" * 16, lo, len(src), lst, cap, grid)
        return sorted(lo + lst[i] for i in range(min(found, cap))), found

    def decode(self, fmt, src, members, out_len, size_only=0, grid=2):
        """members: list of dicts with Member's fields -> (output bytes, [MemberResult])"""
        src = bytes(src)
        arr = (Member * len(members))(*[Member(**m) for m in members])
        res = (MemberResult * len(members))()
        guard = 4096
        dst = C.create_string_buffer(out_len + guard)
        pad = src + b"\0" * 64           # the engine over-allocates its input buffer the same way
        if fmt == FMT_LZ4:
            self.lib.emu_lz4_decompress(pad, dst, arr, res, len(members), grid)
        else:
            self.lib.emu_inflate(fmt, pad, dst, arr, res, len(members), size_only, grid)
        if all(m["dst_off"] + m["dst_cap"] <= out_len for m in members):
            assert dst.raw[out_len:] == b"\0" * guard, "a decoder wrote past the end of its destination"
        return dst.raw[:out_len], list(res)
