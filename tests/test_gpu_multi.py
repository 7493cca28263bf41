"""One process, several GPUs (BASELINE.json configs[3]: chunks sharded across the GPUs of a box via per-GPU submission
queues): with QZB200_DEVICES set, one host-buffer qzCompress call deals its batches over all listed devices and stitches
output and checksum in order.  Needs at least two GPUs (gpurun --gpus 2); skipped otherwise.  The reference hands its
instances out interleaved across devices: src/qatzip.c:795-808, qzGrabInstance :363."""
import ctypes as C
import zlib

import pytest

from harness import qzapi as q

pytestmark = pytest.mark.gpu


def ndev(prod):
    return prod.lib.qzb200DeviceCount()


@pytest.mark.parametrize("fmt", [q.QZ_DEFLATE_GZIP_EXT, q.QZ_DEFLATE_RAW, q.FMT_LZ4])
def test_one_call_over_all_gpus(prod, ref, corpus, monkeypatch, fmt):
    if ndev(prod) < 2:
        pytest.skip("one GPU visible")
    n = 320 << 20                               # ten batches and more: every device gets several
    L = prod.lib
    h_in = L.qzMalloc(n, 0, q.PINNED_MEM)
    cap = L.qzMaxCompressedLength(n, None)
    h_out = L.qzMalloc(cap, 0, q.PINNED_MEM)
    assert h_in and h_out
    corpus.fill(q.Corpus.SILESIA_LIKE, h_in, n)
    # one device first: the bytes to compare with
    sess1 = prod.new_session(fmt=fmt)
    rc, used, made1, crc1 = prod.compress_call(sess1, h_in, n, h_out, cap, 1, crc=0)
    assert rc == q.QZ_OK and used == n and prod.stats(sess1).devices == 1
    one = C.string_at(h_out, made1)
    prod.end_session(sess1)
    monkeypatch.setenv("QZB200_DEVICES", "all")
    sess = prod.new_session(fmt=fmt)
    rc, used, made, crc = prod.compress_call(sess, h_in, n, h_out, cap, 1, crc=0)
    assert rc == q.QZ_OK and used == n
    assert prod.stats(sess).devices == ndev(prod)
    blob = C.string_at(h_out, made)
    assert blob == one                                      # whichever GPU compressed a batch, the stream is the same
    if fmt != q.FMT_LZ4:
        assert crc == crc1 == zlib.crc32(C.string_at(h_in, n))
    # the reference's software path takes it back (RAW: zlib directly)
    if fmt == q.QZ_DEFLATE_RAW:
        assert zlib.decompress(blob, -15) == C.string_at(h_in, n)
    else:
        sr = ref.new_session(fmt=fmt)
        back = (C.c_ubyte * n)()
        rc, used, got = ref.decompress_call(sr, blob, len(blob), C.addressof(back), n)
        assert rc == q.QZ_OK and got == n and C.string_at(C.addressof(back), n) == C.string_at(h_in, n)
        ref.end_session(sr)
    # and our own decompress on a session whose primary device is the next one in turn
    back = L.qzMalloc(n, 0, q.PINNED_MEM)
    sess2 = prod.new_session(fmt=fmt)
    if fmt != q.QZ_DEFLATE_RAW:
        rc, used, got = prod.decompress_call(sess2, h_out, made, back, n)
        assert rc == q.QZ_OK and got == n and C.string_at(back, n) == C.string_at(h_in, n)
    prod.end_session(sess2); prod.end_session(sess)
    for p in (h_in, h_out, back):
        L.qzFree(p)


def test_small_dest_keeps_whole_chunks_across_gpus(prod, corpus, monkeypatch):
    """QZ_BUF_ERROR semantics (reference test mode 17) do not depend on which GPU held the chunk that no longer fits"""
    if ndev(prod) < 2:
        pytest.skip("one GPU visible")
    monkeypatch.setenv("QZB200_DEVICES", "all")
    d = corpus.make(q.Corpus.SILESIA_LIKE, 96 << 20)
    full = prod.compress(d, fmt=q.QZ_DEFLATE_GZIP_EXT)
    sess = prod.new_session(fmt=q.QZ_DEFLATE_GZIP_EXT)
    cap = len(full) // 2
    dst = bytearray(cap)
    rc, used, made = prod.compress_call(sess, d, len(d), dst, cap)
    assert rc == q.QZ_BUF_ERROR and used % 65536 == 0 and 0 < made <= cap
    assert bytes(dst[:made]) == full[:made]
    prod.end_session(sess)
