/* qz_inflate.cu -- sm_100a DEFLATE decoder: several members per warp.
 *
 * Replaces the QAT decompress request at reference src/qatzip.c:2191 (cpaDcDecompressData,
 * stateless, FLUSH_FINAL) and folds in the checks doDecompressOut performs afterwards
 * (reference src/qatzip_utils.c:1483-1532 decompOutCheckSum: checksum vs footer, produced vs
 * ISIZE).  A warp decodes several members at once: one lane per member is its bit-serial Huffman
 * decoder (tables in shared memory, built by all lanes), the whole warp places the decoded
 * tokens.  Output and history live in the caller's destination buffer (HBM / L2), so there is no
 * separate window. */
#include <cuda_runtime.h>
#include <stdint.h>
#include "qz_kernels.cuh"
#include "qz_warp.cuh"
#include "qz_inflate.h"
#include "qz_crc32.h"
#include "qz_adler32.h"

#define FULL 0xffffffffu

#define QZ_INFL_LANE_COPY 16          /* matches up to this long are copied by the lane that owns the token */
struct InflWarpSmem {
    QzInflTables t;
    uint32_t tok[QZ_INFL_BATCH];     /* one decoded batch: lane 0 fills it, every lane places one token */
};

__device__ __forceinline__ uint32_t bcast(uint32_t v) { return __shfl_sync(FULL, v, 0); }
__device__ __forceinline__ uint32_t lanemask_lt() { return qz_lanemask_lt(); }

/* Warp-parallel form of qz_infl_prepare: counts by shared atomics, first codes / offsets by lane 0
 * (15 steps), the (length, symbol) order by taking symbols 32 at a time -- lanes holding equal lengths
 * find each other with match.any and take consecutive places.  Same return values. */
__device__ __noinline__ int warp_infl_prepare(const uint8_t *len, int n, uint16_t *count, uint16_t *first, uint16_t *offs, uint16_t *sorted,
                                              uint32_t *scratch /* 32 words */, uint32_t lane)
{
    uint32_t *cnt = scratch, *fill = scratch + 16;
    if (lane < 16) { cnt[lane] = 0; count[lane] = 0; first[lane] = 0; offs[lane] = 0; }
    __syncwarp();
    for (int s = lane; s < n; s += 32) atomicAdd(&cnt[len[s]], 1u);
    __syncwarp();
    int rc = 0;
    if (lane == 0) {
        int left = 1;
        if ((int)cnt[0] == n) rc = 1;
        else {
            uint32_t code = 0, o = 0;
            first[0] = 0; offs[0] = 0; fill[0] = 0; count[0] = (uint16_t)cnt[0];
            for (int l = 1; l < 16; l++) {
                left <<= 1; left -= (int)cnt[l];
                if (left < 0) { rc = -1; break; }
                count[l] = (uint16_t)cnt[l]; first[l] = (uint16_t)code; offs[l] = (uint16_t)o; fill[l] = o;
                code = (code + cnt[l]) << 1; o += cnt[l];
            }
            if (rc == 0 && left > 0) rc = 1;
        }
    }
    rc = (int)__shfl_sync(FULL, (uint32_t)rc, 0);
    if (rc < 0 || (int)cnt[0] == n) return rc;
    __syncwarp();
    for (int s0 = 0; s0 < n; s0 += 32) {
        const int s = s0 + (int)lane;
        const uint32_t l = s < n ? len[s] : 0u;
        const uint32_t same = __match_any_sync(FULL, l);
        const uint32_t rank = __popc(same & lanemask_lt());
        const uint32_t base = fill[l];
        __syncwarp();
        if (l && rank == 0) fill[l] = base + __popc(same);
        if (l) sorted[base + rank] = (uint16_t)s;
        __syncwarp();
    }
    return rc;
}

/* CRC-32 of dst[0..n) by the whole warp (right-aligned strips + GF(2) tree). */
__device__ uint32_t warp_crc32_global(const uint8_t *p, uint32_t n, const uint32_t *crc_tab, uint32_t lane)
{
    if (n == 0) return 0;
    const uint32_t S = (n + 31) / 32;
    int hi = (int)n - (int)((31 - lane) * S), lo = hi - (int)S;
    if (lo < 0) lo = 0;
    uint32_t c = 0xffffffffu;
    for (int i = lo; i < hi; i++) c = crc_tab[(c ^ p[i]) & 0xff] ^ (c >> 8);
    c = (hi > lo) ? ~c : 0u;
    uint32_t xs = lane < 5 ? qz_crc_xpow8((uint64_t)S << lane) : 0;
#pragma unroll
    for (int lv = 0; lv < 5; lv++) {
        uint32_t other = __shfl_down_sync(FULL, c, 1u << lv);
        uint32_t x = __shfl_sync(FULL, xs, lv);
        if ((lane & ((2u << lv) - 1)) == 0) c = qz_gf2_mul(c, x) ^ other;
    }
    return __shfl_sync(FULL, c, 0);
}

/* CRC-32 of the 4 KiB at p by the whole warp: 128 bytes per lane; xl[lane] = x^(8 * 128 * (31 - lane)) */
#define QZ_INFL_CRC_BLOCK 4096u
__device__ __forceinline__ uint32_t warp_crc32_block(const uint8_t *p, const uint32_t *crc_tab, const uint32_t *xl, uint32_t lane)
{
    const uint4 *q = reinterpret_cast<const uint4 *>(p + lane * 128);
    uint32_t c = 0xffffffffu;
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll 2
        for (int i = 0; i < 8; i++) {
            const uint4 v = q[i];
            const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
            for (int k = 0; k < 4; k++) {
                c = crc_tab[(c ^ w[k]) & 0xff] ^ (c >> 8); c = crc_tab[(c ^ (w[k] >> 8)) & 0xff] ^ (c >> 8);
                c = crc_tab[(c ^ (w[k] >> 16)) & 0xff] ^ (c >> 8); c = crc_tab[(c ^ (w[k] >> 24)) & 0xff] ^ (c >> 8);
            }
        }
    } else {
        const uint8_t *b = p + lane * 128;
        for (int i = 0; i < 128; i++) c = crc_tab[(c ^ b[i]) & 0xff] ^ (c >> 8);
    }
    c = ~c;
    /* crc(block) = xor over lanes of crc(lane's 128 bytes) * x^(8 * 128 * (31 - lane)): one multiply per lane, all at once */
    return __reduce_xor_sync(FULL, lane == 31 ? c : qz_gf2_mul(c, xl[lane]));
}

/* Adler-32 of dst[0..n) by the whole warp: same strips, sums joined as in qz_adler32.h */
__device__ uint32_t warp_adler32_global(const uint8_t *p, uint32_t n, uint32_t lane)
{
    const uint32_t S = n ? (n + 31) / 32 : 1;
    int hi = (int)n - (int)((31 - lane) * S), lo = hi - (int)S;
    if (lo < 0) lo = 0;
    if (hi < lo) hi = lo;
    uint32_t s1, s2;
    qz_adler_block(p + lo, (uint32_t)(hi - lo), &s1, &s2);
#pragma unroll 1
    for (int lv = 0; lv < 5; lv++) {
        const uint32_t o1 = __shfl_down_sync(FULL, s1, 1u << lv), o2 = __shfl_down_sync(FULL, s2, 1u << lv);
        if ((lane & ((2u << lv) - 1)) == 0) qz_adler_join(&s1, &s2, o1, o2, (uint64_t)S << lv);
    }
    return __shfl_sync(FULL, qz_adler_finish(s1, s2, n), 0);
}

/* Thirty-two tokens go to the output, one token per lane (t: this lane's, lanes >= n idle), starting at d[o0]: literals and
 * matches whose source lies wholly before these tokens go out at once, matches that read bytes produced by them follow in
 * order.  Returns the bytes placed.  A distance that reaches before the start of the member's output (d[0]) ends the
 * placing in front of its token and sets *bad_dist: the decode loop leaves that check to this side, which knows every
 * token's position anyway. */
__device__ __forceinline__ uint32_t infl_place(uint8_t *d, uint32_t o0, uint32_t n, uint32_t t, bool wr, uint32_t lane, bool *bad_dist)
{
    bool is_match = lane < n && !qz_tok_is_literal(t);
    uint32_t len = lane < n ? (is_match ? qz_tok_len(t) : 1u) : 0u;
    const uint32_t dist = qz_tok_dist(t);
    uint32_t incl = len;
#pragma unroll
    for (int k = 1; k < 32; k <<= 1) { const uint32_t y = __shfl_up_sync(FULL, incl, k); if (lane >= (uint32_t)k) incl += y; }
    const uint32_t o = o0 + incl - len;                          /* where this lane's token lands */
    uint32_t placed = __shfl_sync(FULL, incl, 31);
    const uint32_t badmask = __ballot_sync(FULL, is_match && dist > o);
    if (badmask) {
        const uint32_t first = __ffs(badmask) - 1;
        placed = __shfl_sync(FULL, o, first) - o0;
        n = first; *bad_dist = true;
        if (lane >= first) { is_match = false; len = 0; }
    }
    const uint32_t span = dist < len ? dist : len;               /* distinct source bytes actually read */
    const bool dep = is_match && (o - dist + span > o0);         /* reads output of this very batch */
    /* short matches that neither overlap themselves nor read this batch: the owning lane copies, all loads first */
    const bool lane_copy = is_match && !dep && len <= QZ_INFL_LANE_COPY && dist >= len;
    if (wr) {
        /* (their sources lie before the batch, so nothing written here is read here: as many steps as the longest of them) */
        const uint32_t steps = __reduce_max_sync(FULL, lane_copy ? len : 0u);
        if (lane < n && !is_match) d[o] = (uint8_t)qz_tok_byte(t);
        const uint8_t *from = d + o - dist;
        for (uint32_t k = 0; k < steps; k += 2) {                /* two bytes a step: half the loop's own instructions */
            const bool c0 = lane_copy && k < len, c1 = lane_copy && k + 1 < len;
            uint8_t v0 = 0, v1 = 0;
            if (c0) v0 = from[k];
            if (c1) v1 = from[k + 1];
            if (c0) d[o + k] = v0;
            if (c1) d[o + k + 1] = v1;
        }
    }
    /* the rest (long, self-overlapping, or fed by this batch) go in order, copied by the whole warp */
    uint32_t depmask = wr ? __ballot_sync(FULL, is_match && !lane_copy) : 0u;
    __syncwarp();
    while (depmask) {
        const uint32_t q = __ffs(depmask) - 1; depmask &= depmask - 1;
        const uint32_t oj = __shfl_sync(FULL, o, q), lj = __shfl_sync(FULL, len, q), dj = __shfl_sync(FULL, dist, q);
        const uint8_t *from = d + oj - dj;
        /* dj < lj: every output byte repeats one of the dj bytes before oj, all already final */
        if (dj >= lj) { for (uint32_t k = lane; k < lj; k += 32) d[oj + k] = from[k]; }
        else { for (uint32_t k = lane; k < lj; k += 32) d[oj + k] = from[k % dj]; }
        __syncwarp();
    }
    return placed;
}

/* The kernel.  A warp works on DPW members at once, one per SLOT: lane s * (32 / DPW) is slot s's DECODER and keeps the
 * member's bit reader, output position and state in its registers; every slot has its own tables, staged input window
 * and token buffer in shared memory.  One round: free slots draw members (largest first); slots between blocks read their
 * block header (the decoder lane reads the code lengths, the warp builds the tables); the warp stages the compressed words
 * the next batch can reach; every decoder turns them into up to 64 tokens (qz_inflate.h: the serial part, one lane);
 * the warp places the tokens 32 at a time, checksums the 4 KiB blocks that have become complete while they are still in
 * cache, and reports finished members.
 * DPW = 1 is what runs (measured: the decode is bound by instruction issue and by the members in flight per SM, which the
 * shared-memory tables fix, so several decoders per warp -- the same instruction stream for all of them, each path of a
 * divergent step paid by everybody -- only trade warps for lanes: 53 / 45 / 29 / 20 GB/s for DPW 1 / 2 / 4 / 8; DESIGN.md
 * section 5).  Four CTAs of eight warps per SM: 64 registers, 6.7 KB of shared memory per member. */
#ifndef QZ_INFL_MIN_CTAS
#define QZ_INFL_MIN_CTAS(dpw) ((dpw) == 1 ? 4 : (dpw) == 2 ? 2 : 1)
#endif
/* The token loop as a function of its own: the compiler then allocates registers for the loop alone instead of sharing them
 * with the kernel's member state.  Window, tables and token buffer are in shared memory, which is said so that the loads stay LDS. */
__device__ __noinline__ int infl_tokens(const uint32_t *inw, uint32_t *lp, const QzInflTables *t, uint32_t *tok, uint32_t *ntok, uint32_t *pos, uint32_t cap)
{
#ifndef QZ_WARP_EMU
    __builtin_assume(__isShared(inw));
    __builtin_assume(__isShared(t));
    __builtin_assume(__isShared(tok));
#endif
    return cap - *pos < QZ_INFL_ROOMY ? qz_inflate_tokens_core<true>(inw, lp, t, tok, ntok, pos, cap)
                                      : qz_inflate_tokens_core<false>(inw, lp, t, tok, ntok, pos, cap);
}

template <int DPW>
__global__ void __launch_bounds__(256, QZ_INFL_MIN_CTAS(DPW)) qzb_inflate_kernel(QzbDecompressJob job)
{
    constexpr uint32_t TL = 32 / DPW;               /* lanes per slot */
    QZ_DYN_SMEM(smem_raw);
    InflWarpSmem *s_w = reinterpret_cast<InflWarpSmem *>(smem_raw);
    __shared__ uint32_t s_crc_tab[256];
    __shared__ uint32_t s_xl[33];           /* [lane] = x^(8 * 128 * (31 - lane)): the block checksum's lane multipliers; [32] = x^(8 * 4096): the fold */
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) s_crc_tab[i] = qz_crc_table_entry(i);
    if (threadIdx.x < 33) s_xl[threadIdx.x] = qz_crc_xpow8(threadIdx.x == 32 ? 4096u : 128u * (31u - threadIdx.x));
    __syncthreads();
    InflWarpSmem *slots = s_w + (size_t)warp * DPW;
    const bool crc_blocks = !job.size_only && job.fmt != QZB_FMT_ZLIB;       /* CRC-32 taken 4 KiB at a time while the output is still in cache */
    const uint32_t myslot = lane / TL;
    const bool is_dec = (lane % TL) == 0;
    const bool wr = !job.size_only;

    /* slot state, meaningful in the slot's decoder lane */
    bool active = false, exhausted = false, in_block = false;
    uint32_t mi = 0, out = 0, status = QZB_ST_OK, bfinal = 0, cap = 0, safe_in = 0, safe_out = 0, crc_run = 0, crc_done = 0;
    const uint8_t *src = job.src; uint8_t *dst = job.dst;
    QzbMember m; m.src_off = 0; m.src_len = 0; m.exact_len = 0; m.dst_off = 0; m.dst_cap = 0; m.exact_out = 0; m.expect_cksum = 0; m.check_cksum = 0;
    QzBitReader br; qz_br_init(&br, job.src, 0);

    for (;;) {
        /* ---- free slots draw members ---- */
        if (is_dec && !active && !exhausted) {
            mi = atomicAdd(job.ticket, 1u);
            if (mi >= job.nmembers) exhausted = true;
            else {
                if (job.order) mi = job.order[mi];       /* the longest members first: the last round of a launch is short ones */
                m = job.members[mi];
                src = job.src + m.src_off; dst = job.dst + m.dst_off; cap = m.dst_cap;
                qz_br_init(&br, src, m.src_len);
                out = 0; status = QZB_ST_OK; bfinal = 0; in_block = false; active = true; safe_in = 0; safe_out = 0; crc_run = 0; crc_done = 0;
            }
        }
        if (__ballot_sync(FULL, is_dec && active) == 0) break;
        bool done = false;                       /* this slot's member ends in this round */

        /* ---- block headers of the slots that are between blocks ---- */
        uint32_t type = 7;
        if (is_dec && active && !in_block) {
            qz_br_refill(&br);
            /* QZ_DEFLATE_RAW chunks that are not the last of the stream end without BFINAL: stop cleanly when nothing
             * but padding is left */
            if (job.fmt == QZB_FMT_RAW && qz_br_exhausted(&br)) { type = 4; done = true; }
            else {
                bfinal = qz_br_bits(&br, 1); type = qz_br_bits(&br, 2);
                if (type == 3) { status = QZB_ST_DATA_ERROR; done = true; }
            }
        }
        /* stored blocks: the decoder checks the lengths, the warp copies */
        uint32_t slen = 0, sstart = 0;
        if (type == 0) {
            const uint32_t drop = br.nacc & 7; br.acc >>= drop; br.nacc -= drop;
            qz_br_refill(&br);
            slen = qz_br_bits(&br, 16);
            const uint32_t nlen = qz_br_bits(&br, 16);
            sstart = qz_br_consumed(&br);
            if ((slen ^ 0xffffu) != nlen) status = QZB_ST_DATA_ERROR;
            else if (slen > br.n || sstart > br.n - slen) status = QZB_ST_IN_TRUNC;
            else if (slen > cap - out) status = QZB_ST_OUT_FULL;
            else qz_br_seek(&br, sstart + slen);
            if (status != QZB_ST_OK) { done = true; slen = 0; }
        }
        uint32_t smask = __ballot_sync(FULL, type == 0 && slen != 0);
        while (smask) {
            const uint32_t j = __ffs(smask) - 1; smask &= smask - 1;
            const uint32_t n = __shfl_sync(FULL, slen, j), st = __shfl_sync(FULL, sstart, j), o = __shfl_sync(FULL, out, j);
            const uint8_t *from = reinterpret_cast<const uint8_t *>(__shfl_sync(FULL, reinterpret_cast<uintptr_t>(src), j)) + st;
            uint8_t *to = reinterpret_cast<uint8_t *>(__shfl_sync(FULL, reinterpret_cast<uintptr_t>(dst), j)) + o;
            if (wr) for (uint32_t i = lane; i < n; i += 32) to[i] = from[i];
        }
        if (type == 0) {
            out += slen;
            if (status == QZB_ST_OK) { safe_in = qz_br_consumed(&br); safe_out = out; if (bfinal) done = true; }      /* byte-aligned behind a stored block */
        }
        __syncwarp();
        /* Huffman blocks: every decoder reads its code lengths (dynamic) or fills in the fixed ones ... */
        uint32_t hlit = 288, hdist = 30;
        if (type == 1) qz_inflate_fixed_lens(&slots[myslot].t);
        else if (type == 2 && qz_inflate_read_dynamic(&br, &slots[myslot].t, &hlit, &hdist) != 0) { status = QZB_ST_DATA_ERROR; done = true; type = 7; }
        __syncwarp();
        /* ... and the whole warp builds the tables, slot after slot (the distance table doubles as scratch until it is
         * cleared) */
        uint32_t hmask = __ballot_sync(FULL, type == 1 || type == 2);
        while (hmask) {
            const uint32_t j = __ffs(hmask) - 1; hmask &= hmask - 1;
            QzInflTables &T = slots[j / TL].t;
            const uint32_t hl = __shfl_sync(FULL, hlit, j), hd = __shfl_sync(FULL, hdist, j);
            const bool bad = warp_infl_prepare(T.lens, (int)hl, T.ll_count, T.ll_first, T.ll_offs, T.ll_sorted, T.d_lut + 64, lane) < 0 ||
                             warp_infl_prepare(T.lens + hl, (int)hd, T.d_count, T.d_first, T.d_offs, T.d_sorted, T.d_lut + 64, lane) < 0;
            if (bad) { if (lane == j) { status = QZB_ST_DATA_ERROR; done = true; } continue; }
            __syncwarp();
            for (uint32_t i = lane; i < (1u << QZ_LL_LUT_BITS); i += 32) T.ll_lut[i] = 0;
            for (uint32_t i = lane; i < (1u << QZ_D_LUT_BITS); i += 32) T.d_lut[i] = 0;
            __syncwarp();
            qz_infl_fill_lut(T.lens, T.ll_count, T.ll_first, T.ll_offs, T.ll_sorted, T.ll_lut, QZ_LL_LUT_BITS, 0, (int)lane, 32);
            qz_infl_fill_lut(T.lens + hl, T.d_count, T.d_first, T.d_offs, T.d_sorted, T.d_lut, QZ_D_LUT_BITS, 1, (int)lane, 32);
            __syncwarp();
            if (lane == j) in_block = true;
        }

        /* ---- every decoder inside a block fills its token buffer (no output touched) ---- */
        int ev = QZI_MATCH; uint32_t ntk = 0, pos = out;
        {
            /* the compressed words the batch can reach go to shared memory first, fetched by the whole warp, slot after slot */
            const bool dec_now = is_dec && active && in_block && !done;
            uint32_t woff = 0, lp = 0;
            if (dec_now) qz_br_where(&br, &woff, &lp);
            uint32_t fmask = __ballot_sync(FULL, dec_now);
            while (fmask) {
                const uint32_t j = __ffs(fmask) - 1; fmask &= fmask - 1;
                const uint8_t *base_j = reinterpret_cast<const uint8_t *>(__shfl_sync(FULL, reinterpret_cast<uintptr_t>(br.base), j));
                const uint32_t end_j = __shfl_sync(FULL, br.end, j), woff_j = __shfl_sync(FULL, woff, j);
                uint32_t *w = slots[j / TL].t.stage;
#pragma unroll
                for (uint32_t i = lane; i < QZ_INFL_INW; i += 32) w[i] = qz_word_at(base_j, end_j, woff_j + 4 * i);
            }
            __syncwarp();
            if (dec_now) {
                ev = infl_tokens(slots[myslot].t.stage, &lp, &slots[myslot].t, slots[myslot].tok, &ntk, &pos, cap);
                qz_br_resume(&br, woff, slots[myslot].t.stage, lp);
                /* past the end of the input the window holds zero bits: see qz_inflate_tokens */
                if (ev == QZI_MATCH && qz_br_overrun(&br)) ev = QZI_ERR_TRUNC;
            }
        }
        __syncwarp();
        /* ---- the batches are placed, slot after slot, by the whole warp: literals and matches whose source lies wholly
         * before the batch go out at once, matches that read bytes produced inside the batch follow in order ---- */
        uint32_t pmask = __ballot_sync(FULL, ntk != 0);
        bool bad_dist = false;
        while (pmask) {
            const uint32_t j = __ffs(pmask) - 1; pmask &= pmask - 1;
            const uint32_t n = __shfl_sync(FULL, ntk, j), o0 = __shfl_sync(FULL, out, j);
            uint8_t *d = reinterpret_cast<uint8_t *>(__shfl_sync(FULL, reinterpret_cast<uintptr_t>(dst), j));
            uint32_t o = o0; bool bad = false;
#pragma unroll 1
            for (uint32_t i0 = 0; i0 < n && !bad; i0 += 32) {
                const uint32_t n32 = min(32u, n - i0);
                o += infl_place(d, o, n32, lane < n32 ? slots[j / TL].tok[i0 + lane] : 0u, wr, lane, &bad);
                __syncwarp();
            }
            if (lane == j) { pos = o; bad_dist = bad; }          /* the output position after the batch, whichever loop decoded it */
        }
        if (is_dec && active && in_block && !done) {
            out = pos;
            if (bad_dist) { status = QZB_ST_DATA_ERROR; done = true; }
            else if (ev == QZI_END_BLOCK) { in_block = false; if (bfinal) done = true; }
            else if (ev == QZI_ERR_DATA) { status = QZB_ST_DATA_ERROR; done = true; }
            else if (ev == QZI_ERR_FULL) { status = QZB_ST_OUT_FULL; done = true; }
            else if (ev == QZI_ERR_TRUNC) { status = QZB_ST_IN_TRUNC; done = true; }
        }

        /* ---- checksum of the 4 KiB blocks that have become complete, while they are still in L1 / L2: the running CRC-32
         * of everything before is multiplied by x^(8 * 4096) and the block's added (zlib's crc32_combine) ---- */
        if (crc_blocks) {
            uint32_t cmask = __ballot_sync(FULL, is_dec && active && status == QZB_ST_OK && out - crc_done >= QZ_INFL_CRC_BLOCK);
            while (cmask) {
                const uint32_t j = __ffs(cmask) - 1;
                const uint32_t from = __shfl_sync(FULL, crc_done, j);
                const uint8_t *d = reinterpret_cast<const uint8_t *>(__shfl_sync(FULL, reinterpret_cast<uintptr_t>(dst), j));
                __syncwarp();
                const uint32_t cb = warp_crc32_block(d + from, s_crc_tab, s_xl, lane);
                if (lane == j) { crc_run = qz_gf2_mul(crc_run, s_xl[32]) ^ cb; crc_done += QZ_INFL_CRC_BLOCK; }
                if (__shfl_sync(FULL, out - crc_done, j) < QZ_INFL_CRC_BLOCK) cmask &= cmask - 1;
            }
        }

        /* ---- finished members: verdict, checksum by the whole warp, result ---- */
        uint32_t consumed = 0;
        if (done) {
            consumed = qz_br_consumed(&br);
            if (status == QZB_ST_OK && qz_br_overrun(&br)) status = QZB_ST_IN_TRUNC;
            if (status == QZB_ST_OK && m.exact_len && consumed != m.src_len) status = QZB_ST_DATA_ERROR;
            if (status == QZB_ST_OK && m.exact_out && out != cap) status = QZB_ST_SIZE;
        }
        uint32_t dmask = __ballot_sync(FULL, done);
        __syncwarp();
        while (dmask) {
            const uint32_t j = __ffs(dmask) - 1; dmask &= dmask - 1;
            const uint32_t stj = __shfl_sync(FULL, status, j), n = __shfl_sync(FULL, out, j);
            const uint8_t *d = reinterpret_cast<const uint8_t *>(__shfl_sync(FULL, reinterpret_cast<uintptr_t>(dst), j));
            uint32_t crc = 0;
            if (stj == QZB_ST_OK && wr) {
                if (job.fmt == QZB_FMT_ZLIB) crc = warp_adler32_global(d, n, lane);
                else {
                    /* what the block-wise pass has not covered yet, joined to the running value */
                    const uint32_t done_j = __shfl_sync(FULL, crc_done, j), run_j = __shfl_sync(FULL, crc_run, j);
                    const uint32_t tail = warp_crc32_global(d + done_j, n - done_j, s_crc_tab, lane);
                    crc = done_j ? (n - done_j ? qz_gf2_mul(run_j, __shfl_sync(FULL, lane == 0 ? qz_crc_xpow8(n - done_j) : 0u, 0)) ^ tail : run_j) : tail;
                }
            }
            if (lane == j) {
                if (status == QZB_ST_OK && wr && m.check_cksum && crc != m.expect_cksum) status = QZB_ST_CKSUM;
                QzbMemberResult r;
                r.status = status; r.consumed = consumed; r.produced = out; r.cksum = crc; r.saw_final = bfinal;
                r.safe_consumed = safe_in; r.safe_produced = safe_out; r.pad = 0;
                job.results[mi] = r;
                active = false;
            }
        }
        __syncwarp();
    }
}

/* Positions in p[lo, n) that look like the start of a gzip member -- the same test as the host's gzip_member_start()
 * (qz_engine.cu; reference: one linear scan per member, src/qatzip_gzip.c:244-261) -- for calls whose compressed bytes are
 * already in device memory: the whole input at HBM speed instead of at the host's.  Offsets relative to lo, unordered. */
__global__ void qzb_gzip_scan_kernel(const uint8_t *p, uint64_t lo, uint64_t n, uint32_t *list, uint32_t cap, uint32_t *count)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q + 10 <= n; q += stride) {
        if (p[q] != 0x1f || p[q + 1] != 0x8b) continue;
        if (p[q + 2] == 8 && (p[q + 3] & 0xe0) == 0 && (p[q + 8] == 0 || p[q + 8] == 2 || p[q + 8] == 4) && (p[q + 9] <= 13 || p[q + 9] == 255)) {
            const uint32_t i = atomicAdd(count, 1u);
            if (i < cap) list[i] = (uint32_t)(q - lo);
        }
    }
}

#ifndef QZ_WARP_EMU
/* decoders per warp (1, 2, 4 or 8); a CTA is 8 warps (4 with eight decoders per warp: 32 slots fill the shared memory) */
static int inflate_cta_warps(int dpw) { return dpw == 8 ? 4 : 8; }
extern "C" size_t qzb_inflate_smem_bytes(int dpw) { return sizeof(InflWarpSmem) * (size_t)inflate_cta_warps(dpw) * (size_t)dpw; }
extern "C" int qzb_inflate_cta_threads(int dpw) { return inflate_cta_warps(dpw) * 32; }
extern "C" int qzb_inflate_cta_slots(int dpw) { return inflate_cta_warps(dpw) * dpw; }
template <int DPW>
static cudaError_t launch_inflate(const QzbDecompressJob &job, int grid, cudaStream_t st)
{
    const size_t smem = qzb_inflate_smem_bytes(DPW);
    cudaError_t e = cudaFuncSetAttribute(qzb_inflate_kernel<DPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    qzb_inflate_kernel<DPW><<<grid, inflate_cta_warps(DPW) * 32, smem, st>>>(job);
    return cudaGetLastError();
}
extern "C" cudaError_t qzb_launch_gzip_scan(const uint8_t *p, uint64_t lo, uint64_t n, uint32_t *list, uint32_t cap, uint32_t *count, int grid, cudaStream_t st)
{
    qzb_gzip_scan_kernel<<<grid, 256, 0, st>>>(p, lo, n, list, cap, count);
    return cudaGetLastError();
}
extern "C" cudaError_t qzb_launch_inflate(const QzbDecompressJob *job, int dpw, int grid, cudaStream_t st)
{
    return dpw == 1 ? launch_inflate<1>(*job, grid, st) : dpw == 2 ? launch_inflate<2>(*job, grid, st) : dpw == 8 ? launch_inflate<8>(*job, grid, st) : launch_inflate<4>(*job, grid, st);
}
#endif
