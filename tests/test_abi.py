"""The C-ABI shared library: loads without a GPU, exports every symbol include/qatzip.h declares,
has the reference's struct layouts, validates parameters exactly like the reference
(reference test/main.c qzSetupParamFuncTest:1114 and the negative tests :3182-3645), and fails
loudly -- QZ_NOSW_NO_HW -- rather than falling back to a CPU path when no CUDA device exists."""
import ctypes as C
import os
import re
import subprocess

import pytest

from harness import qzapi as q
from conftest import has_gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = set()
    for hdr in ("qatzip.h", "qatzip_b200.h"):
        txt = open(os.path.join(ROOT, "include", hdr)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        names |= set(re.findall(r"\b(qz[A-Za-z0-9_]+|qzb200[A-Za-z0-9_]+)\s*\(", txt))
    names -= {"qzLZ4SCallbackFn", "qzAsyncCallbackFn"}
    return sorted(names)


def test_exports_every_declared_symbol(prod):
    names = declared_functions()
    assert len(names) >= 56
    for n in names:
        assert hasattr(prod.lib, n), f"{n} declared in include/*.h but not exported"
    assert hasattr(prod.lib, "logMessage")     # imported by the reference CLI (include/qz_utils.h:119)


def test_no_oracle_or_zlib_in_product():
    """The product must not link the oracle, zlib or lz4: there is no CPU codec behind the ABI."""
    out = subprocess.run(["ldd", q.PRODUCT_SO], capture_output=True, text=True).stdout
    assert "libz" not in out and "lz4" not in out and "oracle" not in out
    syms = subprocess.run(["nm", "-D", "--undefined-only", q.PRODUCT_SO], capture_output=True, text=True).stdout
    for bad in ("inflate", "deflate", "LZ4", "qzo_"):
        assert bad not in syms


def test_struct_layout_matches_reference_abi(tmp_path):
    """sizeof/offsetof of our header == values probed from the reference header."""
    src = tmp_path / "abi.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "qatzip.h"\nint main(void){\n'
                   'printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu ", sizeof(QzSession_T), sizeof(QzStream_T), sizeof(QzSessionParams_T),'
                   'sizeof(QzSessionParamsCommon_T), sizeof(QzSessionParamsDeflate_T), sizeof(QzSessionParamsLZ4_T),'
                   'sizeof(QzSessionParamsDeflateExt_T), sizeof(QzSessionParamsLZ4S_T), sizeof(QzStatus_T));\n'
                   'printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(QzResult_T), offsetof(QzStream_T, pending_in), offsetof(QzStream_T, crc_32),'
                   'offsetof(QzStream_T, reserved), offsetof(QzStream_T, opaque), offsetof(QzSession_T, internal), offsetof(QzSession_T, total_out));\n'
                   'return 0;}\n')
    exe = tmp_path / "abi"
    ours = None
    for inc in (os.path.join(ROOT, "include"), "/root/reference/include"):
        if not os.path.exists(os.path.join(inc, "qatzip.h")):
            continue
        subprocess.run(["gcc", "-I", inc, "-o", str(exe), str(src)], check=True)
        got = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
        if ours is None:
            ours = got
            assert got == ["40", "56", "48", "48", "56", "48", "60", "72", "544", "48", "24", "36", "40", "48", "16", "32"]
        else:
            assert got == ours, "layout differs from the reference header"
    assert C.sizeof(q.QzSession) == 40 and C.sizeof(q.QzStream) == 56


def test_defaults(prod):
    """reference include/qatzip.h:573-600 / src/qatzip.c:97-116"""
    d = q.QzSessionParams()
    assert prod.lib.qzGetDefaults(C.byref(d)) == q.QZ_OK
    assert (d.huffman_hdr, d.direction, d.data_fmt, d.comp_lvl, d.comp_algorithm) == (0, 2, 2, 1, 8)
    assert (d.sw_backup, d.hw_buff_sz, d.strm_buff_sz, d.input_sz_thrshold, d.req_cnt_thrshold, d.wait_cnt_thrshold) == (1, 65536, 65536, 1024, 32, 8)
    l4 = q.QzSessionParamsLZ4()
    assert prod.lib.qzGetDefaultsLZ4(C.byref(l4)) == q.QZ_OK and l4.common_params.comp_algorithm == ord("4")


def test_param_validation(prod):
    """reference test/main.c:1114-1300 (mode 6) and :3182-3645 (modes 13-16)"""
    L = prod.lib

    def setd(**kw):
        d = q.QzSessionParams(); L.qzGetDefaults(C.byref(d))
        for k, v in kw.items():
            setattr(d, k, v)
        return L.qzSetDefaults(C.byref(d))
    assert setd(huffman_hdr=2) == q.QZ_PARAMS
    assert setd(direction=3) == q.QZ_PARAMS
    assert setd(comp_lvl=0) == q.QZ_PARAMS and setd(comp_lvl=10) == q.QZ_PARAMS
    assert setd(sw_backup=2) == q.QZ_PARAMS
    assert setd(hw_buff_sz=0) == q.QZ_PARAMS and setd(hw_buff_sz=1025) == q.QZ_PARAMS
    assert setd(hw_buff_sz=2 * 1024 * 1024) == q.QZ_PARAMS and setd(hw_buff_sz=513 * 1024) == q.QZ_PARAMS and setd(hw_buff_sz=100) == q.QZ_PARAMS
    assert setd(strm_buff_sz=1023) == q.QZ_PARAMS and setd(strm_buff_sz=2 * 1024 * 1024) == q.QZ_PARAMS
    assert setd(input_sz_thrshold=127) == q.QZ_PARAMS
    assert setd(req_cnt_thrshold=0) == q.QZ_PARAMS and setd(req_cnt_thrshold=33) == q.QZ_PARAMS
    assert setd() == q.QZ_OK
    sess = q.QzSession()
    bad = q.QzSessionParamsDeflate(); L.qzGetDefaultsDeflate(C.byref(bad)); bad.common_params.hw_buff_sz = 3000
    assert L.qzSetupSessionDeflate(C.byref(sess), C.byref(bad)) == q.QZ_PARAMS
    bad.common_params.hw_buff_sz = 65536; bad.common_params.comp_lvl = 13
    assert L.qzSetupSessionDeflate(C.byref(sess), C.byref(bad)) == q.QZ_PARAMS
    bad.common_params.comp_lvl = 12; bad.data_fmt = 4
    assert L.qzSetupSessionDeflate(C.byref(sess), C.byref(bad)) == q.QZ_PARAMS
    assert L.qzSetupSession(None, None) == q.QZ_PARAMS
    l4 = q.QzSessionParamsLZ4(); L.qzGetDefaultsLZ4(C.byref(l4)); l4.common_params.comp_algorithm = 8
    assert L.qzSetupSessionLZ4(C.byref(sess), C.byref(l4)) == q.QZ_PARAMS
    assert sess.internal is None


def test_call_argument_checks(prod):
    """reference src/qatzip.c:1853-1861,1883-1891: NULL pointers / last not in {0,1} -> QZ_PARAMS, lengths zeroed"""
    L = prod.lib
    sess = q.QzSession()
    sl, dl = C.c_uint(10), C.c_uint(10)
    buf = C.create_string_buffer(64)
    assert L.qzCompress(C.byref(sess), None, C.byref(sl), buf, C.byref(dl), 1) == q.QZ_PARAMS and sl.value == 0 and dl.value == 0
    sl, dl = C.c_uint(10), C.c_uint(10)
    assert L.qzCompress(C.byref(sess), buf, C.byref(sl), buf, C.byref(dl), 2) == q.QZ_PARAMS and sl.value == 0
    assert L.qzDecompress(C.byref(sess), buf, None, buf, C.byref(dl)) == q.QZ_PARAMS
    assert L.qzCompressStream(C.byref(sess), None, 1) == q.QZ_PARAMS
    st = q.QzStream(); st.in_sz = 5; st.out_sz = 5
    assert L.qzCompressStream(C.byref(sess), C.byref(st), 3) == q.QZ_PARAMS and st.in_sz == 0 and st.out_sz == 0
    assert L.qzEndStream(C.byref(sess), None) == q.QZ_PARAMS


def test_max_compressed_length(prod):
    """reference src/qatzip.c:3022-3068, probed values in SURVEY.md section 8b"""
    f = prod.lib.qzMaxCompressedLength
    assert f(0, None) == 34 and f(65536, None) == 73808 and f(512 << 20, None) == 603979856 and f(0xFFFFFFFF, None) == 0


def test_memory_api_without_gpu_falls_back_like_reference(prod):
    """qzMalloc(COMMON) always yields memory; qzMemFindAddr says whether it is pinned."""
    L = prod.lib
    p = L.qzMalloc(100000, 0, q.COMMON_MEM)
    assert p
    assert L.qzMemFindAddr(p) in (0, 1)
    L.qzFree(p)
    if not has_gpu():
        assert not L.qzMalloc(4096, 0, q.PINNED_MEM)     # reference src/qatzip_mem.c:211-215: pinned request fails -> NULL


@pytest.mark.skipif(has_gpu(), reason="this asserts the no-device behaviour")
def test_no_device_fails_loudly(prod):
    L = prod.lib
    sess = q.QzSession()
    assert L.qzInit(C.byref(sess), 1) == q.QZ_NOSW_NO_HW
    assert L.qzSetupSessionDeflate(C.byref(sess), None) == q.QZ_NOSW_NO_HW
    src, dst = C.create_string_buffer(b"x" * 4096, 4096), C.create_string_buffer(8192)
    sl, dl = C.c_uint(4096), C.c_uint(8192)
    assert L.qzCompress(C.byref(sess), src, C.byref(sl), dst, C.byref(dl), 1) == q.QZ_NOSW_NO_HW
    assert sl.value == 0 and dl.value == 0
    sl, dl = C.c_uint(4096), C.c_uint(8192)
    assert L.qzDecompress(C.byref(sess), src, C.byref(sl), dst, C.byref(dl)) == q.QZ_NOSW_NO_HW
