"""Sharding of one logical buffer across ranks (one process per GPU), SURVEY.md section 8e.

Chunks are independent (reference: CPA_DC_STATELESS, src/qatzip_utils.c:332; the reference spreads
them over QAT instances of different devices, src/qatzip.c:795-808), so a buffer is cut into
`world` contiguous, chunk-aligned ranges; every rank runs the ordinary C-ABI calls on its range
with its own per-GPU submission queue.  Nothing on the data path crosses ranks.  The only
exchange is control-plane: each rank's output length and CRC (a few bytes), from which every
rank knows where its bytes sit in the concatenated stream and what the whole-buffer CRC is.
"""


def shard_range(total_bytes, world, rank, chunk):
    """Contiguous, chunk-aligned [lo, hi) of `total_bytes` for `rank`; earlier ranks take the extra chunks."""
    nchunks = (total_bytes + chunk - 1) // chunk if total_bytes else 0
    base, extra = divmod(nchunks, world)
    first = rank * base + min(rank, extra)
    count = base + (1 if rank < extra else 0)
    lo = min(first * chunk, total_bytes)
    hi = min((first + count) * chunk, total_bytes)
    return lo, hi


def _gf2_mul(a, b):
    p = 0
    for _ in range(32):
        if a & 0x80000000:
            p ^= b
        a = (a << 1) & 0xFFFFFFFF
        b = (b >> 1) ^ (0xEDB88320 if b & 1 else 0)
    return p


def crc32_combine(crc_a, crc_b, len_b):
    """crc(A||B) from crc(A), crc(B), len(B) -- same arithmetic as qz_crc32.h / zlib crc32_combine."""
    xp, sq, n = 0x80000000, 0x00800000, len_b
    while n:
        if n & 1:
            xp = _gf2_mul(xp, sq)
        sq = _gf2_mul(sq, sq)
        n >>= 1
    return _gf2_mul(crc_a, xp) ^ crc_b


def exchange_layout(dist, local_in_len, local_out_len, local_crc):
    """all_gather of three integers per rank.  Returns (out_offset_of_this_rank, total_out, whole_crc)."""
    import torch
    mine = torch.tensor([local_in_len, local_out_len, local_crc], dtype=torch.int64)
    if dist is None or not dist.is_initialized():
        return 0, local_out_len, local_crc
    if dist.get_backend() == "nccl":
        mine = mine.cuda()
    rows = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(rows, mine)
    rows = [[int(v) for v in r.tolist()] for r in rows]
    rank = dist.get_rank()
    off = sum(r[1] for r in rows[:rank])
    crc = 0
    for r in rows:
        if r[0] == 0:
            continue
        crc = r[2] if crc == 0 else crc32_combine(crc, r[2], r[0])
    return off, sum(r[1] for r in rows), crc
