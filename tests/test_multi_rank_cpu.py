"""The N>1 path on CPU: two gloo ranks shard one buffer, each compresses its range through the same
C-ABI-shaped call (the oracle port stands in for the GPU here -- the sharding logic is codec
agnostic), exchange only lengths and CRCs, and the concatenation must be one valid stream."""
import os
import sys
import zlib

import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port_file, total, fmt, q_out):
    import torch
    import torch.distributed as dist
    from harness import qzapi as q
    from qatzip_b200 import shard
    dist.init_process_group("gloo", init_method=f"file://{port_file}", rank=rank, world_size=world)
    data = q.Corpus().make(q.Corpus.SILESIA_LIKE, total)          # every rank can regenerate any range
    lo, hi = shard.shard_range(total, world, rank, 65536)
    port = q.OraclePort()
    blob, crc = port.compress(data[lo:hi], fmt, want_crc=True) if hi > lo else (b"", 0)
    off, total_out, whole_crc = shard.exchange_layout(dist, hi - lo, len(blob), crc)
    # test-only: collect the bytes on rank 0 to check the concatenation (the product never does this)
    sizes = [None] * world
    dist.all_gather_object(sizes, (off, len(blob)))
    blobs = [None] * world
    dist.all_gather_object(blobs, blob)
    if rank == 0:
        cat = b"".join(blobs)
        assert [s[0] for s in sizes] == [sum(len(b) for b in blobs[:r]) for r in range(world)]
        assert total_out == len(cat)
        assert port.decompress(cat, fmt, total + 8) == data
        if fmt != q.FMT_LZ4:
            assert whole_crc == zlib.crc32(data)
        q_out.put("ok")
    dist.barrier()
    dist.destroy_process_group()


def _run(total, fmt, world=2, tmp="/tmp"):
    ctx = mp.get_context("spawn")
    qo = ctx.Queue()
    rv = os.path.join(tmp, f"qz_gloo_{os.getpid()}_{total}_{fmt}")
    if os.path.exists(rv):
        os.remove(rv)
    ps = [ctx.Process(target=_worker, args=(r, world, rv, total, fmt, qo)) for r in range(world)]
    [p.start() for p in ps]
    [p.join(120) for p in ps]
    assert all(p.exitcode == 0 for p in ps), [p.exitcode for p in ps]
    assert qo.get(timeout=5) == "ok"


def test_shard_ranges():
    from qatzip_b200 import shard
    for total in (0, 1, 65535, 65536, 65537, 1 << 20, (1 << 20) + 5, 16 << 30):
        for world in (1, 2, 3, 4, 8):
            rs = [shard.shard_range(total, world, r, 65536) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == total
            for (a, b), (c, d) in zip(rs, rs[1:]):
                assert b == c and a <= b
            assert all(lo % 65536 == 0 for lo, hi in rs if hi > lo)
            assert max(hi - lo for lo, hi in rs) - min(hi - lo for lo, hi in rs) <= 65536 + 65535
    assert shard.crc32_combine(zlib.crc32(b"hello "), zlib.crc32(b"world"), 5) == zlib.crc32(b"hello world")


def test_two_ranks_gzip_ext(tmp_path):
    from harness import qzapi as q
    _run((3 << 20) + 12345, q.QZ_DEFLATE_GZIP_EXT, tmp=str(tmp_path))


def test_two_ranks_lz4_and_tiny(tmp_path):
    from harness import qzapi as q
    _run(70000, q.FMT_LZ4, tmp=str(tmp_path))          # rank 1 gets a 4464-byte tail
    _run(1000, q.QZ_DEFLATE_GZIP, tmp=str(tmp_path))   # rank 1 gets nothing
