/* qz_inflate.cu -- sm_100a DEFLATE decoder: one warp per member.
 *
 * Replaces the QAT decompress request at reference src/qatzip.c:2191 (cpaDcDecompressData,
 * stateless, FLUSH_FINAL) and folds in the checks doDecompressOut performs afterwards
 * (reference src/qatzip_utils.c:1483-1532 decompOutCheckSum: checksum vs footer, produced vs
 * ISIZE).  Lane 0 is the bit-serial Huffman decoder (tables in the warp's shared-memory
 * slice, built by all lanes); literals are stored as they are decoded and every
 * back-reference / stored block is copied by the whole warp.  Output and history live in
 * the caller's destination buffer (HBM / L2), so there is no separate window. */
#include <cuda_runtime.h>
#include <stdint.h>
#include "qz_kernels.cuh"
#include "qz_inflate.h"
#include "qz_crc32.h"
#include "qz_adler32.h"

#define FULL 0xffffffffu

#define QZ_INFL_BATCH 32
struct InflWarpSmem {
    QzInflTables t;
    uint16_t code_of[320];
    uint32_t tok[QZ_INFL_BATCH];     /* one decoded batch: lane 0 fills it, every lane places one token */
};

__device__ __forceinline__ uint32_t bcast(uint32_t v) { return __shfl_sync(FULL, v, 0); }

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

__global__ void __launch_bounds__(256, 4) qzb_inflate_kernel(QzbDecompressJob job)
{
    __shared__ InflWarpSmem s_w[8];
    __shared__ uint32_t s_crc_tab[256];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) s_crc_tab[i] = qz_crc_table_entry(i);
    __syncthreads();
    InflWarpSmem &ws = s_w[warp];
    QzInflTables &T = ws.t;

    for (;;) {
        uint32_t mi = 0;
        if (lane == 0) mi = atomicAdd(job.ticket, 1u);
        mi = bcast(mi);
        if (mi >= job.nmembers) break;
        const QzbMember m = job.members[mi];
        const uint8_t *src = job.src + m.src_off;
        uint8_t *dst = job.dst + m.dst_off;
        const uint32_t cap = m.dst_cap;
        QzBitReader br;
        qz_br_init(&br, src, m.src_len);
        uint32_t out = 0, status = QZB_ST_OK, bfinal = 0;
        const bool wr = !job.size_only;

        while (!bfinal && status == QZB_ST_OK) {
            uint32_t type = 0;
            if (lane == 0) {
                qz_br_refill(&br);
                /* QZ_DEFLATE_RAW chunks that are not the last of the stream end without BFINAL:
                 * stop cleanly when nothing but padding is left */
                if (job.fmt == QZB_FMT_RAW && qz_br_consumed(&br) >= br.n && br.phantom * 8 >= br.nacc) type = 4;
                else { bfinal = qz_br_bits(&br, 1); type = qz_br_bits(&br, 2); }
            }
            type = bcast(type); bfinal = bcast(bfinal);
            if (type == 4) break;
            if (type == 3) { status = QZB_ST_DATA_ERROR; break; }
            if (type == 0) {
                uint32_t len = 0, start = 0, st = QZB_ST_OK;
                if (lane == 0) {
                    uint32_t drop = br.nacc & 7; br.acc >>= drop; br.nacc -= drop;
                    qz_br_refill(&br);
                    len = qz_br_bits(&br, 16);
                    uint32_t nlen = qz_br_bits(&br, 16);
                    start = qz_br_consumed(&br);
                    if ((len ^ 0xffffu) != nlen) st = QZB_ST_DATA_ERROR;
                    else if (start + len > br.n) st = QZB_ST_IN_TRUNC;
                    else if (out + len > cap) st = QZB_ST_OUT_FULL;
                    else { br.pos = start + len; br.acc = 0; br.nacc = 0; br.phantom = 0; }
                }
                st = bcast(st); len = bcast(len); start = bcast(start);
                if (st != QZB_ST_OK) { status = st; break; }
                if (wr) for (uint32_t i = lane; i < len; i += 32) dst[out + i] = src[start + i];
                out += len;
                __syncwarp();
                continue;
            }
            /* Huffman block: lane 0 reads the code lengths, the warp builds the tables */
            uint32_t st = QZB_ST_OK, hlit = 288, hdist = 30;
            if (lane == 0) {
                if (type == 1) qz_inflate_fixed_lens(&T);
                else if (qz_inflate_read_dynamic(&br, &T, &hlit, &hdist) != 0) st = QZB_ST_DATA_ERROR;
                if (st == QZB_ST_OK) {
                    if (qz_infl_prepare(T.lens, (int)hlit, T.ll_count, T.ll_sorted, ws.code_of) < 0) st = QZB_ST_DATA_ERROR;
                    else if (qz_infl_prepare(T.lens + hlit, (int)hdist, T.d_count, T.d_sorted, ws.code_of + 288) < 0) st = QZB_ST_DATA_ERROR;
                }
            }
            st = bcast(st); hlit = bcast(hlit); hdist = bcast(hdist);
            if (st != QZB_ST_OK) { status = st; break; }
            for (uint32_t i = lane; i < (1u << QZ_LL_LUT_BITS); i += 32) T.ll_lut[i] = 0;
            for (uint32_t i = lane; i < (1u << QZ_D_LUT_BITS); i += 32) T.d_lut[i] = 0;
            __syncwarp();
            qz_infl_fill_lut(T.lens, ws.code_of, (int)hlit, T.ll_lut, QZ_LL_LUT_BITS, (int)lane, 32);
            qz_infl_fill_lut(T.lens + hlit, ws.code_of + 288, (int)hdist, T.d_lut, QZ_D_LUT_BITS, (int)lane, 32);
            __syncwarp();
            /* lane 0 decodes a batch of tokens (no output touched); then every lane places one:
             * literals and matches whose source lies wholly before the batch go out at once, matches
             * that read bytes produced inside the batch follow in order, copied by the whole warp */
            for (;;) {
                int ev = 0; uint32_t ntk = 0, pos = out;
                if (lane == 0) ev = qz_inflate_tokens(&br, &T, ws.tok, QZ_INFL_BATCH, &ntk, &pos, cap);
                __syncwarp();
                ev = (int)bcast((uint32_t)ev); ntk = bcast(ntk);
                const uint32_t t = lane < ntk ? ws.tok[lane] : 0u;
                const bool is_match = lane < ntk && (t >> 31);
                const uint32_t len = lane < ntk ? (is_match ? ((t >> 16) & 0xff) + 3 : 1u) : 0u;
                const uint32_t dist = (t & 0x7fff) + 1;
                uint32_t incl = len;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(FULL, incl, o); if (lane >= (uint32_t)o) incl += y; }
                const uint32_t o = out + incl - len;                       /* where this lane's token lands */
                const uint32_t span = dist < len ? dist : len;               /* distinct source bytes actually read */
                const bool dep = is_match && (o - dist + span > out);        /* reads output of this very batch */
                if (lane < ntk && wr) {
                    if (!is_match) dst[o] = (uint8_t)t;
                    else if (!dep) {
                        const uint8_t *from = dst + o - dist;
                        if (dist >= len) { for (uint32_t k = 0; k < len; k++) dst[o + k] = __ldcg(from + k); }
                        else { for (uint32_t k = 0; k < len; k++) dst[o + k] = __ldcg(from + k % dist); }
                    }
                }
                uint32_t depmask = wr ? __ballot_sync(FULL, dep) : 0u;
                __syncwarp();
                while (depmask) {
                    const uint32_t j = __ffs(depmask) - 1; depmask &= depmask - 1;
                    const uint32_t oj = __shfl_sync(FULL, o, j), lj = __shfl_sync(FULL, len, j), dj = __shfl_sync(FULL, dist, j);
                    const uint8_t *from = dst + oj - dj;
                    /* dj < lj: every output byte repeats one of the dj bytes before oj, all already final */
                    if (dj >= lj) { for (uint32_t k = lane; k < lj; k += 32) dst[oj + k] = __ldcg(from + k); }
                    else { for (uint32_t k = lane; k < lj; k += 32) dst[oj + k] = __ldcg(from + k % dj); }
                    __syncwarp();
                }
                out += __shfl_sync(FULL, incl, 31);
                if (ev == QZI_END_BLOCK) break;
                if (ev == QZI_ERR_DATA) { status = QZB_ST_DATA_ERROR; break; }
                if (ev == QZI_ERR_FULL) { status = QZB_ST_OUT_FULL; break; }
                if (ev == QZI_ERR_TRUNC) { status = QZB_ST_IN_TRUNC; break; }
            }
        }
        /* verdict */
        uint32_t consumed = bcast(lane == 0 ? qz_br_consumed(&br) : 0u);
        uint32_t over = bcast(lane == 0 ? (uint32_t)qz_br_overrun(&br) : 0u);
        if (status == QZB_ST_OK && over) status = QZB_ST_IN_TRUNC;
        if (status == QZB_ST_OK && m.exact_len && consumed != m.src_len) status = QZB_ST_DATA_ERROR;
        if (status == QZB_ST_OK && m.exact_out && out != cap) status = QZB_ST_SIZE;
        uint32_t crc = 0;
        if (status == QZB_ST_OK && wr) {
            __syncwarp();
            crc = (job.fmt == QZB_FMT_ZLIB) ? warp_adler32_global(dst, out, lane) : warp_crc32_global(dst, out, s_crc_tab, lane);
            if (m.check_cksum && crc != m.expect_cksum) status = QZB_ST_CKSUM;
        }
        if (lane == 0) {
            QzbMemberResult r;
            r.status = status; r.consumed = consumed; r.produced = out; r.cksum = crc; r.saw_final = bfinal;
            r.pad[0] = r.pad[1] = r.pad[2] = 0;
            job.results[mi] = r;
        }
        __syncwarp();
    }
}

/* ------------------------------------------------------------------------------------------
 * Lane-per-member decoder.  Huffman decoding is serial inside a member, so the warp-per-member
 * kernel above keeps 31 of 32 lanes idle while lane 0 decodes.  When a batch holds many members
 * whose output positions are known (gzip / gzip-ext), each LANE takes its own member instead:
 * private bit reader in registers, private decode tables in shared memory (odd word stride, so
 * the 32 lanes of a warp start in 32 different banks), literals stored directly, matches copied
 * 8 bytes per L2 round trip.  The same host+device decode routines are used (qz_inflate.h).
 * CRC-32 verification runs afterwards in qzb_crc_verify_kernel with a whole warp per member. */
struct InflLaneSmem {
    QzInflTables t;
    uint16_t code_of[320];
    uint32_t pad;                /* sizeof % 8 == 4: consecutive lanes' tables start one bank apart */
};
static_assert((sizeof(InflLaneSmem) / 4) % 2 == 1, "lane table stride must be an odd number of words");

/* Every lane runs this flat state machine; one trip through the loop does a bounded amount of work
 * (one symbol, or 8 bytes of a match, or 16 bytes of a stored block), so no lane ever waits for
 * another lane's long inner loop -- the only long state is the (rare) block header. */
enum { LS_FETCH = 0, LS_HEADER, LS_DECODE, LS_COPY, LS_STORED, LS_FINISH, LS_EXIT };

__global__ void __launch_bounds__(64) qzb_inflate_lanes_kernel(QzbDecompressJob job)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    InflLaneSmem &ws = reinterpret_cast<InflLaneSmem *>(smem_raw)[threadIdx.x];
    QzInflTables &T = ws.t;
    const uint32_t stride = gridDim.x * blockDim.x;
    uint32_t mi = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t state = LS_FETCH, status = QZB_ST_OK, bfinal = 0, out = 0, cap = 0, mlen = 0, mdist = 0, s_src = 0, s_rem = 0;
    uint32_t exact_len = 0, exact_out = 0, src_len = 0;
    const uint8_t *src = nullptr; uint8_t *dst = nullptr;
    QzBitReader br; qz_br_init(&br, nullptr, 0);

    /* all lanes stay in the loop until the whole warp is done and re-converge after every trip:
     * without the explicit __syncwarp the 32 lanes drift apart and run one after another */
    while (__any_sync(0xffffffffu, state != LS_EXIT)) {
        if (state == LS_EXIT) { /* idle */ }
        else if (state == LS_DECODE) {
            qz_br_refill(&br);
            const uint32_t e = T.ll_lut[br.acc & ((1u << QZ_LL_LUT_BITS) - 1)];
            int sym;
            if (e) { sym = (int)(e >> 4); br.acc >>= (e & 15); br.nacc -= (e & 15); }
            else sym = qz_infl_slow(&br, T.ll_count, T.ll_sorted);
            if (sym < 0) { status = QZB_ST_DATA_ERROR; state = LS_FINISH; }
            else if (sym < 256) {
                if (out >= cap) { status = QZB_ST_OUT_FULL; state = LS_FINISH; }
                else dst[out++] = (uint8_t)sym;
            } else if (sym == 256) state = bfinal ? LS_FINISH : LS_HEADER;
            else {
                sym -= 257;
                uint32_t eb = 0, len = sym < 29 ? qz_len_base((uint32_t)sym, &eb) : 0;
                if (br.nacc < 48) qz_br_refill(&br);
                len += qz_br_bits(&br, eb);
                const uint32_t de = T.d_lut[br.acc & ((1u << QZ_D_LUT_BITS) - 1)];
                int ds;
                if (de) { ds = (int)(de >> 4); br.acc >>= (de & 15); br.nacc -= (de & 15); }
                else ds = qz_infl_slow(&br, T.d_count, T.d_sorted);
                uint32_t dist = 0;
                if (ds >= 0 && ds < 30) { dist = qz_dist_base((uint32_t)ds, &eb); dist += qz_br_bits(&br, eb); }
                if (sym >= 29 || dist == 0 || dist > out) { status = QZB_ST_DATA_ERROR; state = LS_FINISH; }
                else if (out + len > cap) { status = QZB_ST_OUT_FULL; state = LS_FINISH; }
                else { mlen = len; mdist = dist; state = LS_COPY; }
            }
        } else if (state == LS_COPY) {
            /* dst[out + j] = dst[out + j - mdist] for up to 8 bytes; all loads first (one L2 round trip) */
            const uint8_t *from = dst + out - mdist;
            const uint32_t m = mlen < 8 ? mlen : 8, span = mdist < 8 ? mdist : 8;
            uint8_t b0 = 0, b1 = 0, b2 = 0, b3 = 0, b4 = 0, b5 = 0, b6 = 0, b7 = 0;
            if (0 < span) b0 = __ldcg(from + 0);
            if (1 < span) b1 = __ldcg(from + 1);
            if (2 < span) b2 = __ldcg(from + 2);
            if (3 < span) b3 = __ldcg(from + 3);
            if (4 < span) b4 = __ldcg(from + 4);
            if (5 < span) b5 = __ldcg(from + 5);
            if (6 < span) b6 = __ldcg(from + 6);
            if (7 < span) b7 = __ldcg(from + 7);
            if (mdist < 8) {      /* short period: later bytes repeat earlier ones */
                if (mdist == 1) { b1 = b0; b2 = b0; b3 = b0; b4 = b0; b5 = b0; b6 = b0; b7 = b0; }
                else if (mdist == 2) { b2 = b0; b3 = b1; b4 = b0; b5 = b1; b6 = b0; b7 = b1; }
                else if (mdist == 3) { b3 = b0; b4 = b1; b5 = b2; b6 = b0; b7 = b1; }
                else if (mdist == 4) { b4 = b0; b5 = b1; b6 = b2; b7 = b3; }
                else if (mdist == 5) { b5 = b0; b6 = b1; b7 = b2; }
                else if (mdist == 6) { b6 = b0; b7 = b1; }
                else { b7 = b0; }
            }
            uint8_t *o = dst + out;
            if (0 < m) o[0] = b0;
            if (1 < m) o[1] = b1;
            if (2 < m) o[2] = b2;
            if (3 < m) o[3] = b3;
            if (4 < m) o[4] = b4;
            if (5 < m) o[5] = b5;
            if (6 < m) o[6] = b6;
            if (7 < m) o[7] = b7;
            out += m; mlen -= m;
            if (mlen == 0) state = LS_DECODE;
        } else if (state == LS_STORED) {
            const uint32_t m = s_rem < 16 ? s_rem : 16;
            for (uint32_t j = 0; j < m; j++) dst[out + j] = src[s_src + j];
            out += m; s_src += m; s_rem -= m;
            if (s_rem == 0) { br.pos = s_src; br.acc = 0; br.nacc = 0; br.phantom = 0; state = bfinal ? LS_FINISH : LS_HEADER; }
        } else if (state == LS_HEADER) {
            qz_br_refill(&br);
            bfinal = qz_br_bits(&br, 1);
            const uint32_t type = qz_br_bits(&br, 2);
            if (type == 3) { status = QZB_ST_DATA_ERROR; state = LS_FINISH; }
            else if (type == 0) {
                const uint32_t drop = br.nacc & 7; br.acc >>= drop; br.nacc -= drop;
                qz_br_refill(&br);
                const uint32_t len = qz_br_bits(&br, 16), nlen = qz_br_bits(&br, 16), start = qz_br_consumed(&br);
                if ((len ^ 0xffffu) != nlen) { status = QZB_ST_DATA_ERROR; state = LS_FINISH; }
                else if (start + len > br.n) { status = QZB_ST_IN_TRUNC; state = LS_FINISH; }
                else if (out + len > cap) { status = QZB_ST_OUT_FULL; state = LS_FINISH; }
                else if (len == 0) { br.pos = start; br.acc = 0; br.nacc = 0; br.phantom = 0; state = bfinal ? LS_FINISH : LS_HEADER; }
                else { s_src = start; s_rem = len; state = LS_STORED; }
            } else {
                uint32_t hlit = 288, hdist = 30; bool ok = true;
                if (type == 1) qz_inflate_fixed_lens(&T);
                else ok = qz_inflate_read_dynamic(&br, &T, &hlit, &hdist) == 0;
                ok = ok && qz_infl_prepare(T.lens, (int)hlit, T.ll_count, T.ll_sorted, ws.code_of) >= 0
                        && qz_infl_prepare(T.lens + hlit, (int)hdist, T.d_count, T.d_sorted, ws.code_of + 288) >= 0;
                if (!ok) { status = QZB_ST_DATA_ERROR; state = LS_FINISH; }
                else {
                    for (uint32_t i = 0; i < (1u << QZ_LL_LUT_BITS); i++) T.ll_lut[i] = 0;
                    for (uint32_t i = 0; i < (1u << QZ_D_LUT_BITS); i++) T.d_lut[i] = 0;
                    qz_infl_fill_lut(T.lens, ws.code_of, (int)hlit, T.ll_lut, QZ_LL_LUT_BITS, 0, 1);
                    qz_infl_fill_lut(T.lens + hlit, ws.code_of + 288, (int)hdist, T.d_lut, QZ_D_LUT_BITS, 0, 1);
                    state = LS_DECODE;
                }
            }
        } else if (state == LS_FINISH) {
            const uint32_t consumed = qz_br_consumed(&br);
            if (status == QZB_ST_OK && qz_br_overrun(&br)) status = QZB_ST_IN_TRUNC;
            if (status == QZB_ST_OK && (exact_len & 1) && consumed != src_len) status = QZB_ST_DATA_ERROR;
            if (status == QZB_ST_OK && exact_out && out != cap) status = QZB_ST_SIZE;
            QzbMemberResult r;
            r.status = status; r.consumed = consumed; r.produced = out; r.cksum = 0; r.saw_final = bfinal;
            r.pad[0] = r.pad[1] = r.pad[2] = 0;
            job.results[mi] = r;
            mi += stride;
            state = LS_FETCH;
        } else {   /* LS_FETCH */
            if (mi >= job.nmembers) state = LS_EXIT;
            else {
                const QzbMember m = job.members[mi];
                src = job.src + m.src_off; dst = job.dst + m.dst_off; cap = m.dst_cap; src_len = m.src_len;
                exact_len = m.exact_len; exact_out = m.exact_out;
                qz_br_init(&br, src, m.src_len);
                out = 0; status = QZB_ST_OK; bfinal = 0; state = LS_HEADER;
            }
        }
        __syncwarp();
    }
}

/* one warp per member: CRC-32 of what the lane decoder produced, compared with the footer */
__global__ void __launch_bounds__(256) qzb_crc_verify_kernel(QzbDecompressJob job)
{
    __shared__ uint32_t s_crc_tab[256];
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) s_crc_tab[i] = qz_crc_table_entry(i);
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t mi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; mi < job.nmembers; mi += nw) {
        const QzbMember m = job.members[mi];
        QzbMemberResult r = job.results[mi];
        if (r.status != QZB_ST_OK) continue;
        const uint32_t crc = warp_crc32_global(job.dst + m.dst_off, r.produced, s_crc_tab, lane);
        if (lane == 0) {
            r.cksum = crc;
            if (m.check_cksum && crc != m.expect_cksum) r.status = QZB_ST_CKSUM;
            job.results[mi] = r;
        }
    }
}

extern "C" cudaError_t qzb_launch_inflate_lanes(const QzbDecompressJob *job, int sm_count, cudaStream_t st)
{
    const int threads = 64;
    const size_t smem = sizeof(InflLaneSmem) * threads;
    cudaError_t e = cudaFuncSetAttribute(qzb_inflate_lanes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int grid = (int)((job->nmembers + threads - 1) / threads);
    if (grid > sm_count * 2) grid = sm_count * 2;
    qzb_inflate_lanes_kernel<<<grid, threads, smem, st>>>(*job);
    int cgrid = (int)((job->nmembers + 7) / 8);
    if (cgrid > sm_count * 8) cgrid = sm_count * 8;
    qzb_crc_verify_kernel<<<cgrid, 256, 0, st>>>(*job);
    return cudaGetLastError();
}

extern "C" cudaError_t qzb_launch_inflate(const QzbDecompressJob *job, int grid, cudaStream_t st)
{
    qzb_inflate_kernel<<<grid, 256, 0, st>>>(*job);
    return cudaGetLastError();
}
