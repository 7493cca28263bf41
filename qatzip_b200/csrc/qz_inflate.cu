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
#include "qz_warp.cuh"
#include "qz_inflate.h"
#include "qz_crc32.h"
#include "qz_adler32.h"

#define FULL 0xffffffffu

#define QZ_INFL_BATCH 32
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
    QZ_DYN_SMEM(smem_raw);
    InflWarpSmem *s_w = reinterpret_cast<InflWarpSmem *>(smem_raw);
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
                if (job.fmt == QZB_FMT_RAW && qz_br_exhausted(&br)) type = 4;
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
                    else qz_br_seek(&br, start + len);
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
            }
            st = bcast(st); hlit = bcast(hlit); hdist = bcast(hdist);
            if (st != QZB_ST_OK) { status = st; break; }
            __syncwarp();
            /* the whole warp builds the tables (the literal/length table's tail doubles as scratch until it is cleared) */
            if (warp_infl_prepare(T.lens, (int)hlit, T.ll_count, T.ll_first, T.ll_offs, T.ll_sorted, T.ll_lut + 512, lane) < 0 ||
                warp_infl_prepare(T.lens + hlit, (int)hdist, T.d_count, T.d_first, T.d_offs, T.d_sorted, T.ll_lut + 512, lane) < 0) { status = QZB_ST_DATA_ERROR; break; }
            __syncwarp();
            for (uint32_t i = lane; i < (1u << QZ_LL_LUT_BITS); i += 32) T.ll_lut[i] = 0;
            for (uint32_t i = lane; i < (1u << QZ_D_LUT_BITS); i += 32) T.d_lut[i] = 0;
            __syncwarp();
            qz_infl_fill_lut(T.lens, T.ll_count, T.ll_first, T.ll_offs, T.ll_sorted, T.ll_lut, QZ_LL_LUT_BITS, 0, (int)lane, 32);
            qz_infl_fill_lut(T.lens + hlit, T.d_count, T.d_first, T.d_offs, T.d_sorted, T.d_lut, QZ_D_LUT_BITS, 1, (int)lane, 32);
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
                const bool is_match = lane < ntk && !qz_tok_is_literal(t);
                const uint32_t len = lane < ntk ? (is_match ? qz_tok_len(t) : 1u) : 0u;
                const uint32_t dist = qz_tok_dist(t);
                uint32_t incl = len;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(FULL, incl, o); if (lane >= (uint32_t)o) incl += y; }
                const uint32_t o = out + incl - len;                       /* where this lane's token lands */
                const uint32_t span = dist < len ? dist : len;               /* distinct source bytes actually read */
                const bool dep = is_match && (o - dist + span > out);        /* reads output of this very batch */
                /* short matches that neither overlap themselves nor read this batch: the owning lane copies,
                 * all loads first (ordinary cached loads: what this warp wrote is in this SM's L1 or in L2) */
                const bool lane_copy = is_match && !dep && len <= QZ_INFL_LANE_COPY && dist >= len;
                if (lane < ntk && wr) {
                    if (!is_match) dst[o] = (uint8_t)qz_tok_byte(t);
                    else if (lane_copy) {
                        const uint8_t *from = dst + o - dist;
                        uint8_t v[QZ_INFL_LANE_COPY];
#pragma unroll
                        for (uint32_t k = 0; k < QZ_INFL_LANE_COPY; k++) v[k] = k < len ? from[k] : (uint8_t)0;
#pragma unroll
                        for (uint32_t k = 0; k < QZ_INFL_LANE_COPY; k++) if (k < len) dst[o + k] = v[k];
                    }
                }
                /* the rest (long, self-overlapping, or fed by this batch) go in order, copied by the whole warp */
                uint32_t depmask = wr ? __ballot_sync(FULL, is_match && !lane_copy) : 0u;
                __syncwarp();
                while (depmask) {
                    const uint32_t j = __ffs(depmask) - 1; depmask &= depmask - 1;
                    const uint32_t oj = __shfl_sync(FULL, o, j), lj = __shfl_sync(FULL, len, j), dj = __shfl_sync(FULL, dist, j);
                    const uint8_t *from = dst + oj - dj;
                    /* dj < lj: every output byte repeats one of the dj bytes before oj, all already final */
                    if (dj >= lj) { for (uint32_t k = lane; k < lj; k += 32) dst[oj + k] = from[k]; }
                    else { for (uint32_t k = lane; k < lj; k += 32) dst[oj + k] = from[k % dj]; }
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

#ifndef QZ_WARP_EMU
extern "C" cudaError_t qzb_launch_inflate(const QzbDecompressJob *job, int grid, cudaStream_t st)
{
    const size_t smem = sizeof(InflWarpSmem) * 8;          /* 8 warps per CTA, 4 CTAs per SM */
    cudaError_t e = cudaFuncSetAttribute(qzb_inflate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    qzb_inflate_kernel<<<grid, 256, smem, st>>>(*job);
    return cudaGetLastError();
}
#endif
