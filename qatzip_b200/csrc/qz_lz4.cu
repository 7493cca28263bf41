/* qz_lz4.cu -- sm_100a LZ4 block codec + xxHash32, one warp per piece (compress) / frame (decompress).
 *
 * Replaces the QAT LZ4 session (reference src/qatzip_utils.c:292-298: compType CPA_DC_LZ4,
 * 64 KiB max block, no block checksum, XXH32 content checksum returned in res.checksum) used by
 * the submit calls at reference src/qatzip.c:1542 / :2191.  The frame around the blocks is
 * written by qzb_frame_kernel (qz_deflate.cu), layout from reference src/qatzip_lz4.c:104-143.
 *
 * Compress: every PIECE of a chunk becomes one LZ4 block (4-byte LE size header, bit 31 =
 * stored), so a 64 KiB chunk is a frame of 8 independent blocks.  Match search is the same
 * 32-positions-per-step hash probe as the deflate kernel, under the LZ4 end-of-block rules
 * (last match starts >= 12 bytes before the end, last 5 bytes are literals).
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include "qz_kernels.cuh"
#include "qz_warp.cuh"
#include "qz_match.cuh"
#include "qz_xxh32.h"

#define FULL 0xffffffffu
#define LZ4_HB 11
#define LZ4_LANE_CAP 36

__device__ __forceinline__ uint32_t lz_lane() { return threadIdx.x & 31; }
__device__ __forceinline__ uint32_t lz_lt() { return qz_lanemask_lt(); }
__device__ __forceinline__ uint32_t lz_ld32u(const uint8_t *base, uint32_t off)
{
    const uint32_t *w = reinterpret_cast<const uint32_t *>(base) + (off >> 2);
    return __funnelshift_r(w[0], w[1], (off & 3) * 8);
}

template <int PIECE_LOG2>
struct Lz4WarpSmem {
    uint8_t piece[(1 << PIECE_LOG2) + 32];
    uint16_t table[1 << LZ4_HB];
};

/* bytes of the "length >= 15" extension: 0 if v < 15, else 1 + (v - 15) / 255 */
__device__ __forceinline__ uint32_t lz4_ext(uint32_t v) { return v < 15 ? 0u : 1u + (v - 15) / 255u; }
__device__ __forceinline__ uint8_t *lz4_put_ext(uint8_t *o, uint32_t v)
{
    if (v >= 15) { v -= 15; while (v >= 255) { *o++ = 255; v -= 255; } *o++ = (uint8_t)v; }
    return o;
}

template <int PIECE_LOG2>
__global__ void __launch_bounds__(512) qzb_lz4_pieces_kernel(QzbCompressJob job)
{
    constexpr int PIECE = 1 << PIECE_LOG2;
    typedef Lz4WarpSmem<PIECE_LOG2> WS;
    QZ_DYN_SMEM(smem_raw);
    const uint32_t lane = lz_lane(), warp = threadIdx.x >> 5;
    WS &ws = reinterpret_cast<WS *>(smem_raw)[warp];
    const uint32_t gwarp = blockIdx.x * (blockDim.x >> 5) + warp;
    /* match records: two words each (pos | len << 16, dist) */
    uint32_t *recs = job.tok_scratch + (size_t)gwarp * QZB_TOK_STRIDE(PIECE);

    for (;;) {
        uint32_t g = 0;
        if (lane == 0) g = atomicAdd(job.ticket, 1u);
        g = __shfl_sync(FULL, g, 0);
        if (g >= job.npieces) break;
        const uint32_t chunk = g / job.pieces_per_chunk, k = g - chunk * job.pieces_per_chunk;
        const uint64_t chunk_off = (uint64_t)chunk * job.chunk_sz;
        const uint64_t rem = job.src_len > chunk_off ? job.src_len - chunk_off : 0;
        const uint32_t chunk_len = rem < job.chunk_sz ? (uint32_t)rem : job.chunk_sz;
        const uint32_t p_off = k << PIECE_LOG2;
        const uint32_t n = chunk_len > p_off ? min((uint32_t)PIECE, chunk_len - p_off) : 0u;
        const uint8_t *src = job.src + chunk_off + p_off;
        uint8_t *slot = job.slots + (size_t)g * job.slot_stride;
        if (n == 0) { if (lane == 0) job.piece_len[g] = 0; continue; }     /* empty frame: no block at all */

        /* ---- load ---- */
        {
            uint4 *d4 = reinterpret_cast<uint4 *>(ws.piece);
            if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
                const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
                uint32_t nv = n >> 4;
                for (uint32_t i = lane; i < nv; i += 32) d4[i] = __ldg(s4 + i);
                for (uint32_t i = (nv << 4) + lane; i < n; i += 32) ws.piece[i] = src[i];
            } else for (uint32_t i = lane; i < n; i += 32) ws.piece[i] = src[i];
            ws.piece[n + lane] = 0;
            for (uint32_t i = lane; i < (1u << LZ4_HB) / 2; i += 32) reinterpret_cast<uint32_t *>(ws.table)[i] = 0xffffffffu;
            __syncwarp();
        }
        /* ---- match + select: records only for matches ---- */
        uint32_t nrec = 0;
        {
            const uint32_t mstart_lim = n >= 12 ? n - 12 : 0;      /* last match must start <= n-12 */
            const uint32_t mend_lim = n >= 5 ? n - 5 : 0;           /* and end <= n-5 */
            uint32_t entry = 0;
            for (uint32_t base = 0; base < n; base += 32) {
                const uint32_t p = base + lane;
                const uint32_t v = lz_ld32u(ws.piece, p);
                const bool can = n >= 13 && p <= mstart_lim;
                const uint32_t h = (v * 2654435761u) >> (32 - LZ4_HB);
                uint32_t cand = can ? ws.table[h] : 0xffffu;
                __syncwarp();
                if (can) ws.table[h] = (uint16_t)p;
                __syncwarp();
                uint32_t L = 0;
                const uint32_t maxl = can ? mend_lim - p : 0;
                if (cand != 0xffffu && lz_ld32u(ws.piece, cand) == v) {
                    uint32_t l = 4;
                    while (l < LZ4_LANE_CAP) {
                        uint32_t x = lz_ld32u(ws.piece, p + l) ^ lz_ld32u(ws.piece, cand + l);
                        if (x) { l += (__ffs(x) - 1) >> 3; break; }
                        l += 4;
                    }
                    L = min(l, maxl);
                }
                const uint32_t M = __ballot_sync(FULL, L >= 4);
                uint32_t matchmask = 0, cur = entry;
                if (cur >= 32) { entry = cur - 32; continue; }
                for (;;) {
                    uint32_t rest = M & (FULL << cur);
                    if (!rest) { cur = 32; break; }
                    uint32_t m = __ffs(rest) - 1;
                    matchmask |= 1u << m;
                    uint32_t Lm = __shfl_sync(FULL, L, m);
                    if (Lm >= LZ4_LANE_CAP) {
                        const uint32_t pm = base + m, cm = __shfl_sync(FULL, cand, m), mx = mend_lim - pm;
                        Lm = LZ4_LANE_CAP;
                        while (Lm < mx) {
                            uint32_t kk = Lm + lane;
                            bool eq = kk < mx && ws.piece[pm + kk] == ws.piece[cm + kk];
                            uint32_t bal = __ballot_sync(FULL, eq);
                            if (bal == FULL) { Lm += 32; continue; }
                            Lm += __ffs(~bal) - 1; break;
                        }
                        Lm = min(Lm, mx);
                        if (lane == m) L = Lm;
                    }
                    cur = m + Lm;
                    if (cur >= 32) break;
                }
                entry = cur - 32;
                if ((matchmask >> lane) & 1) {
                    uint32_t r = nrec + __popc(matchmask & lz_lt());
                    recs[2 * r] = p | (L << 16);
                    recs[2 * r + 1] = p - cand;
                }
                nrec += __popc(matchmask);
            }
        }
        __syncwarp();
        /* ---- size pass: sequence sizes -> offsets ---- */
        uint8_t *out = slot + 4;
        uint32_t total = 0, prev_end_carry = 0;
        for (uint32_t r0 = 0; r0 < nrec; r0 += 32) {
            uint32_t r = r0 + lane, pos = 0, len = 0, dist = 0, sz = 0, lit = 0, lit_start = 0;
            const bool have = r < nrec;
            if (have) { uint32_t a = __ldcg(recs + 2 * r); dist = __ldcg(recs + 2 * r + 1); pos = a & 0xffff; len = a >> 16; }
            uint32_t my_end = have ? pos + len : 0;
            uint32_t prev_end = __shfl_up_sync(FULL, my_end, 1);
            if (lane == 0) prev_end = prev_end_carry;
            if (have) { lit_start = prev_end; lit = pos - lit_start; sz = 1 + lz4_ext(lit) + lit + 2 + lz4_ext(len - 4); }
            uint32_t incl = sz;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(FULL, incl, o); if (lane >= (uint32_t)o) incl += y; }
            if (have && total + incl <= n) {        /* anything that would not beat "stored" is never written */
                uint8_t *o = out + total + incl - sz;
                *o++ = (uint8_t)((lit >= 15 ? 15u : lit) << 4 | (len - 4 >= 15 ? 15u : len - 4));
                o = lz4_put_ext(o, lit);
                for (uint32_t i = 0; i < lit; i++) o[i] = ws.piece[lit_start + i];
                o += lit;
                *o++ = (uint8_t)dist; *o++ = (uint8_t)(dist >> 8);
                lz4_put_ext(o, len - 4);
            }
            total += __shfl_sync(FULL, incl, 31);
            const uint32_t last_lane = min(31u, nrec - r0 - 1);
            prev_end_carry = __shfl_sync(FULL, my_end, last_lane);
        }
        /* last literals */
        const uint32_t tail = n - prev_end_carry;
        const uint32_t tail_sz = 1 + lz4_ext(tail) + tail;
        const bool stored = total + tail_sz >= n;
        uint32_t blk;
        if (!stored) {
            uint8_t *o = out + total;
            if (lane == 0) { *o = (uint8_t)((tail >= 15 ? 15u : tail) << 4); lz4_put_ext(o + 1, tail); }
            o += 1 + lz4_ext(tail);
            for (uint32_t i = lane; i < tail; i += 32) o[i] = ws.piece[prev_end_carry + i];
            blk = total + tail_sz;
            if (lane == 0) { slot[0] = (uint8_t)blk; slot[1] = (uint8_t)(blk >> 8); slot[2] = (uint8_t)(blk >> 16); slot[3] = 0; }
        } else {
            __syncwarp();
            for (uint32_t i = lane; i < n; i += 32) out[i] = ws.piece[i];
            blk = n;
            if (lane == 0) { slot[0] = (uint8_t)n; slot[1] = (uint8_t)(n >> 8); slot[2] = (uint8_t)(n >> 16); slot[3] = 0x80; }
        }
        if (lane == 0) job.piece_len[g] = blk + 4;
        __syncwarp();
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Window kernel (hw_buff_sz >= 64 KiB): 64 KiB of a chunk become ONE LZ4 block, as in the reference's sessions
 * (lz4BlockMaxSize = 64 KiB, reference src/qatzip_utils.c:292-298, src/qatzip_sw.c:443-471).  A CTA holds one window in
 * shared memory next to its warps' hash tables that take the rest of it (about 5000 entries each); sixteen warps match a
 * sub-piece of 4 KiB each with the stages of qz_match.cuh (every position has the whole window in front of it as history),
 * leave (position, length, distance) records in the L2 scratch, size their sequences, scan the sizes through shared
 * memory and write the block at byte offsets; literals are read from global memory again, so the next window's TMA bulk
 * copy (issued by warp 0 as soon as everybody has matched) overlaps the encoding.  A block that would not be smaller than
 * its input is stored. */
#define QZL_MAX_WARPS 16u
struct Lz4WindowShared { uint32_t ticket[2]; uint32_t nrec[QZL_MAX_WARPS], last_end[QZL_MAX_WARPS], bytes[QZL_MAX_WARPS]; };
/* bytes of a warp's sub-piece when `nw` warps share the window: whole tiles, a multiple of 16 (16 warps: 4096, 12 warps: 5504) */
__host__ __device__ __forceinline__ uint32_t lz4_sub_bytes(uint32_t nw) { return ((65536u + nw - 1) / nw + 31u) & ~31u; }
__host__ __device__ __forceinline__ uint32_t lz4_table_stride(uint32_t tent) { return (tent + 8u) & ~7u; }
__host__ __device__ __forceinline__ uint32_t lz4_window_smem(uint32_t tent, uint32_t nw) { return QZM_FRONT_PAD + 65536u + QZM_TAIL_PAD + nw * 2u * lz4_table_stride(tent); }

/* Sequences of the records [0, nrec): sizes, and with `out` their bytes at out[0...).  prev_end = where the literals of the
 * first sequence start.  Returns the bytes of all sequences; *last_end = end of the last match (prev_end if there is none). */
__device__ __forceinline__ uint32_t lz4_encode_records(const uint32_t *recs, uint32_t nrec, uint32_t prev_end, uint8_t *out, const uint8_t *src, uint32_t lane, uint32_t *last_end)
{
    uint32_t total = 0, carry = prev_end;
    for (uint32_t r0 = 0; r0 < nrec; r0 += 32) {
        const uint32_t r = r0 + lane;
        uint32_t pos = 0, len = 0, dist = 0, sz = 0, lit = 0, lit_start = 0;
        const bool have = r < nrec;
        if (have) { const uint32_t a = __ldcg(recs + 2 * r); dist = __ldcg(recs + 2 * r + 1); pos = a & 0xffff; len = a >> 16; }
        const uint32_t my_end = have ? pos + len : 0;
        uint32_t before = __shfl_up_sync(FULL, my_end, 1);
        if (lane == 0) before = carry;
        if (have) { lit_start = before; lit = pos - lit_start; sz = 1 + lz4_ext(lit) + lit + 2 + lz4_ext(len - 4); }
        uint32_t incl = sz;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(FULL, incl, o); if (lane >= (uint32_t)o) incl += y; }
        if (out) {
            uint8_t *o = out + total + incl - sz;
            if (have) {
                *o++ = (uint8_t)((lit >= 15 ? 15u : lit) << 4 | (len - 4 >= 15 ? 15u : len - 4));
                o = lz4_put_ext(o, lit);
                if (lit <= 16) for (uint32_t i = 0; i < lit; i++) o[i] = src[lit_start + i];
                uint8_t *q = o + lit;
                *q++ = (uint8_t)dist; *q++ = (uint8_t)(dist >> 8);
                lz4_put_ext(q, len - 4);
            }
            /* long literal runs are copied by the whole warp */
            uint32_t big = __ballot_sync(FULL, have && lit > 16);
            while (big) {
                const uint32_t l = __ffs(big) - 1; big &= big - 1;
                const uint32_t n = __shfl_sync(FULL, lit, l), from = __shfl_sync(FULL, lit_start, l);
                uint8_t *to = reinterpret_cast<uint8_t *>(__shfl_sync(FULL, reinterpret_cast<uintptr_t>(o), l));
                for (uint32_t i = lane; i < n; i += 32) to[i] = src[from + i];
            }
        }
        total += __shfl_sync(FULL, incl, 31);
        carry = __shfl_sync(FULL, my_end, min(31u, nrec - r0 - 1));
    }
    *last_end = carry;
    return total;
}

template <int QZL_WARPS>
__global__ void __launch_bounds__(QZL_WARPS * 32) qzb_lz4_window_kernel(QzbCompressJob job)
{
    constexpr uint32_t QZL_SUB = ((65536u + QZL_WARPS - 1) / QZL_WARPS + 31u) & ~31u;
    QZ_DYN_SMEM(smem_raw);
    __shared__ uint64_t s_mbar[1];
    __shared__ Lz4WindowShared G;
    const uint32_t lane = lz_lane(), wg = threadIdx.x >> 5;
    const uint32_t tent = job.tent, tstride = lz4_table_stride(tent);
    uint8_t *win = smem_raw + QZM_FRONT_PAD;
    uint16_t *tables = reinterpret_cast<uint16_t *>(smem_raw + QZM_FRONT_PAD + 65536u + QZM_TAIL_PAD);
    uint16_t *table = tables + (size_t)wg * tstride;
    uint32_t *recs = job.tok_scratch + (size_t)(blockIdx.x * QZL_WARPS + wg) * QZB_TOK_STRIDE(QZL_SUB);
    const uint32_t wpc = job.pieces_per_chunk / 8;          /* windows per chunk */
    if (threadIdx.x == 0) qz_mbar_init(&s_mbar[0]);
    __syncthreads();

    /* warp 0: draw window k and have it copied in (everything behind the 16-byte-aligned part by hand) */
    auto fetch = [&](uint32_t k) {
        uint32_t tk = 0;
        if (lane == 0) tk = atomicAdd(job.ticket, 1u);
        tk = __shfl_sync(FULL, tk, 0);
        uint32_t bulk = 0;
        const uint8_t *src = job.src;
        if (tk < job.ngroups) {
            const uint32_t chunk = tk / wpc, blk = tk - chunk * wpc;
            const uint64_t off = (uint64_t)chunk * job.chunk_sz + (uint64_t)blk * 65536u;
            const uint32_t wlen = (uint32_t)min((uint64_t)65536u, min((uint64_t)chunk * job.chunk_sz + job.chunk_sz, job.src_len) - off);
            src = job.src + off;
            bulk = (reinterpret_cast<uintptr_t>(src) & 15) == 0 ? wlen & ~15u : 0u;
            for (uint32_t i = bulk + lane; i < wlen; i += 32) win[i] = src[i];
            for (uint32_t i = lane; i < QZM_TAIL_PAD; i += 32) win[wlen + i] = 0;
        }
        __syncwarp();
        if (lane == 0) { G.ticket[k & 1] = tk; qz_bulk_load_arrive(win, src, bulk, &s_mbar[0]); }
        __syncwarp();
    };
    if (wg == 0) fetch(0);
    for (uint32_t k = 0;; k++) {
        qz_mbar_wait(&s_mbar[0], k);
        const uint32_t gi = G.ticket[k & 1];
        if (gi >= job.ngroups) break;
        const uint32_t chunk = gi / wpc, blk = gi - chunk * wpc;
        const uint32_t g0 = chunk * job.pieces_per_chunk + blk * 8;
        const uint64_t off = (uint64_t)chunk * job.chunk_sz + (uint64_t)blk * 65536u;
        const uint32_t wlen = (uint32_t)min((uint64_t)65536u, min((uint64_t)chunk * job.chunk_sz + job.chunk_sz, job.src_len) - off);
        const uint8_t *src = job.src + off;
        const uint32_t nsub = (wlen + QZL_SUB - 1) / QZL_SUB, npc = (wlen + 8191u) >> 13;
        const uint32_t p0 = wg * QZL_SUB, n = wlen > p0 ? min(QZL_SUB, wlen - p0) : 0u;
        if (n) qzm_prepass<1>(win, p0 + n, p0, p0 + n, table, tent, lane);
        __syncthreads();
        qzm_seed_tables<1>(tables, tstride, nsub, tent, threadIdx.x, QZL_WARPS * 32);
        __syncthreads();
        QzmLz4Sink sink = { recs, 0 };
        if (n) qzm_match_piece<1>(win, wlen, p0, p0 + n, table, tent, sink, lane);
        uint32_t my_last = 0;
        if (sink.nrec) { const uint32_t a = __ldcg(recs + 2 * (sink.nrec - 1)); my_last = (a & 0xffff) + (a >> 16); }
        if (lane == 0) { G.nrec[wg] = sink.nrec; G.last_end[wg] = my_last; }
        __syncthreads();
        /* everybody has matched: the next window may come in while this one is encoded */
        if (wg == 0) fetch(k + 1);
        /* where this warp's first literal run starts: the end of the last match in front of its records */
        uint32_t prev_end = 0, lastw = 0;
        for (uint32_t i = 0; i < QZL_WARPS; i++) { if (G.nrec[i]) { lastw = i; if (i < wg) prev_end = G.last_end[i]; } }
        uint32_t dummy;
        uint32_t bytes = lz4_encode_records(recs, sink.nrec, prev_end, nullptr, src, lane, &dummy);
        const uint32_t final_end = G.last_end[lastw];            /* 0 when the window has no match at all */
        const uint32_t tail = wlen - final_end, tail_sz = 1 + lz4_ext(tail) + tail;
        if (wg == lastw) bytes += tail_sz;
        if (lane == 0) G.bytes[wg] = bytes;
        __syncthreads();
        uint32_t before = 0, total = 0;
        for (uint32_t i = 0; i < QZL_WARPS; i++) { const uint32_t bi = G.bytes[i]; if (i < wg) before += bi; total += bi; }
        uint8_t *slot = job.slots + (size_t)g0 * job.slot_stride, *out = slot + 4;
        uint32_t blk_bytes;
        if (total < wlen) {
            uint32_t le;
            const uint32_t mine = lz4_encode_records(recs, sink.nrec, prev_end, out + before, src, lane, &le);
            if (wg == lastw) {
                uint8_t *o = out + before + mine;
                if (lane == 0) { *o = (uint8_t)((tail >= 15 ? 15u : tail) << 4); lz4_put_ext(o + 1, tail); }
                o += 1 + lz4_ext(tail);
                for (uint32_t i = lane; i < tail; i += 32) o[i] = src[final_end + i];
            }
            blk_bytes = total;
            if (threadIdx.x == 0) { slot[0] = (uint8_t)total; slot[1] = (uint8_t)(total >> 8); slot[2] = (uint8_t)(total >> 16); slot[3] = 0; }
        } else {
            for (uint32_t i = threadIdx.x; i < wlen; i += QZL_WARPS * 32) out[i] = src[i];
            blk_bytes = wlen;
            if (threadIdx.x == 0) { slot[0] = (uint8_t)wlen; slot[1] = (uint8_t)(wlen >> 8); slot[2] = (uint8_t)(wlen >> 16); slot[3] = 0x80; }
        }
        if (lane == 0 && wg < npc) job.piece_len[g0 + wg] = wg == 0 ? blk_bytes + 4 : 0u;
        __syncthreads();            /* G is reused by the next window */
    }
}

/* ---- XXH32 of every chunk: 4 lanes per chunk (one per accumulator), 8 chunks per warp ----
 * reference src/xxhash.c:404-437 (stripe loop) / :328-400 (finalize) */
__global__ void __launch_bounds__(256) qzb_xxh32_chunks_kernel(QzbCompressJob job)
{
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, c = tid >> 2, q = tid & 3, lane = threadIdx.x & 31;
    const bool live = c < job.nchunks;
    const uint64_t chunk_off = (uint64_t)c * job.chunk_sz;
    const uint64_t rem = (live && job.src_len > chunk_off) ? job.src_len - chunk_off : 0;
    const uint32_t len = rem < job.chunk_sz ? (uint32_t)rem : job.chunk_sz;
    const uint8_t *p = job.src + chunk_off;
    const bool aligned = (reinterpret_cast<uintptr_t>(p) & 3) == 0;
    uint32_t v = q == 0 ? QZ_XP1 + QZ_XP2 : q == 1 ? QZ_XP2 : q == 2 ? 0u : 0u - QZ_XP1;
    const uint32_t nstripes = len >> 4;
    if (aligned) { const uint32_t *w = reinterpret_cast<const uint32_t *>(p) + q; for (uint32_t s = 0; s < nstripes; s++) v = qz_xxh_round(v, __ldg(w + 4 * s)); }
    else for (uint32_t s = 0; s < nstripes; s++) v = qz_xxh_round(v, qz_xxh_rd32(p + 16 * s + 4 * q));
    const uint32_t gb = lane & ~3u;
    const uint32_t v1 = __shfl_sync(FULL, v, gb), v2 = __shfl_sync(FULL, v, gb + 1), v3 = __shfl_sync(FULL, v, gb + 2), v4 = __shfl_sync(FULL, v, gb + 3);
    if (live && q == 0) {
        uint32_t h = len >= 16 ? qz_rotl32(v1, 1) + qz_rotl32(v2, 7) + qz_rotl32(v3, 12) + qz_rotl32(v4, 18) : QZ_XP5;
        h += len;
        const uint8_t *t = p + (nstripes << 4), *end = p + len;
        while (t + 4 <= end) { h = qz_rotl32(h + qz_xxh_rd32(t) * QZ_XP3, 17) * QZ_XP4; t += 4; }
        while (t < end) { h = qz_rotl32(h + (*t++) * QZ_XP5, 11) * QZ_XP1; }
        job.chunk_cksum[c] = qz_xxh_avalanche(h);
    }
}

/* ---- decompress: one warp per frame ---- */
__device__ __forceinline__ uint32_t lz_rd32g(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }

/* XXH32 of dst[0..n) by 4 lanes of the warp (result valid in all lanes) */
__device__ uint32_t warp_xxh32_global(const uint8_t *p, uint32_t n, uint32_t lane)
{
    const uint32_t q = lane & 3;
    uint32_t v = q == 0 ? QZ_XP1 + QZ_XP2 : q == 1 ? QZ_XP2 : q == 2 ? 0u : 0u - QZ_XP1;
    const uint32_t nstripes = n >> 4;
    if (lane < 4) for (uint32_t s = 0; s < nstripes; s++) v = qz_xxh_round(v, lz_rd32g(p + 16 * s + 4 * q));
    const uint32_t v1 = __shfl_sync(FULL, v, 0), v2 = __shfl_sync(FULL, v, 1), v3 = __shfl_sync(FULL, v, 2), v4 = __shfl_sync(FULL, v, 3);
    uint32_t h = 0;
    if (lane == 0) {
        h = n >= 16 ? qz_rotl32(v1, 1) + qz_rotl32(v2, 7) + qz_rotl32(v3, 12) + qz_rotl32(v4, 18) : QZ_XP5;
        h += n;
        const uint8_t *t = p + (nstripes << 4), *end = p + n;
        while (t + 4 <= end) { h = qz_rotl32(h + lz_rd32g(t) * QZ_XP3, 17) * QZ_XP4; t += 4; }
        while (t < end) { h = qz_rotl32(h + (*t++) * QZ_XP5, 11) * QZ_XP1; }
        h = qz_xxh_avalanche(h);
    }
    return __shfl_sync(FULL, h, 0);
}

__global__ void __launch_bounds__(256) qzb_lz4_decompress_kernel(QzbDecompressJob job)
{
    const uint32_t lane = threadIdx.x & 31;
    for (;;) {
        uint32_t mi = 0;
        if (lane == 0) mi = atomicAdd(job.ticket, 1u);
        mi = __shfl_sync(FULL, mi, 0);
        if (mi >= job.nmembers) break;
        const QzbMember m = job.members[mi];
        const uint8_t *src = job.src + m.src_off;
        uint8_t *dst = job.dst + m.dst_off;
        const uint32_t cap = m.dst_cap, n = m.src_len;
        const bool blk_cksum = (m.exact_len & 2) != 0;
        uint32_t ip = 0, op = 0, status = QZB_ST_OK;
        /* blocks until the payload is used up (the EndMark sits after src_len) */
        while (ip < n && status == QZB_ST_OK) {
            if (ip + 4 > n) { status = QZB_ST_IN_TRUNC; break; }
            const uint32_t bh = lz_rd32g(src + ip); ip += 4;
            const uint32_t bs = bh & 0x7fffffffu;
            if (bs > n - ip) { status = QZB_ST_IN_TRUNC; break; }
            if (bh & 0x80000000u) {
                if (bs > cap - op) { status = QZB_ST_OUT_FULL; break; }
                for (uint32_t i = lane; i < bs; i += 32) dst[op + i] = src[ip + i];
                op += bs; ip += bs;
                __syncwarp();
            } else {
                const uint32_t bend = ip + bs;
                while (ip < bend && status == QZB_ST_OK) {
                    /* lane 0 parses one sequence header, the warp copies */
                    uint32_t lit = 0, mlen = 0, off = 0, lit_src = 0, nip = 0, st = QZB_ST_OK;
                    if (lane == 0) {
                        uint32_t i = ip;
                        const uint32_t tok = src[i++];
                        lit = tok >> 4;
                        if (lit == 15) { uint32_t s; do { if (i >= bend) { st = QZB_ST_DATA_ERROR; break; } s = src[i++]; lit += s; } while (s == 255); }
                        lit_src = i;
                        if (st == QZB_ST_OK && lit > bend - i) st = QZB_ST_DATA_ERROR;
                        i += lit;
                        if (st == QZB_ST_OK && i < bend) {
                            if (i + 2 > bend) st = QZB_ST_DATA_ERROR;
                            else {
                                off = src[i] | (uint32_t)src[i + 1] << 8; i += 2;
                                mlen = tok & 15;
                                if (mlen == 15) { uint32_t s; do { if (i >= bend) { st = QZB_ST_DATA_ERROR; break; } s = src[i++]; mlen += s; } while (s == 255); }
                                mlen += 4;
                                if (off == 0) st = QZB_ST_DATA_ERROR;
                            }
                        }
                        nip = i;
                    }
                    st = __shfl_sync(FULL, st, 0); lit = __shfl_sync(FULL, lit, 0); mlen = __shfl_sync(FULL, mlen, 0);
                    off = __shfl_sync(FULL, off, 0); lit_src = __shfl_sync(FULL, lit_src, 0); nip = __shfl_sync(FULL, nip, 0);
                    if (st != QZB_ST_OK) { status = st; break; }
                    if (lit > cap - op || mlen > cap - op - lit) { status = QZB_ST_OUT_FULL; break; }
                    for (uint32_t i = lane; i < lit; i += 32) dst[op + i] = src[lit_src + i];
                    op += lit;
                    __syncwarp();
                    if (mlen) {
                        if (off > op) { status = QZB_ST_DATA_ERROR; break; }
                        const uint8_t *from = dst + op - off;
                        if (off >= mlen) { for (uint32_t i = lane; i < mlen; i += 32) dst[op + i] = from[i]; }
                        else { for (uint32_t i = lane; i < mlen; i += 32) dst[op + i] = from[i % off]; }
                        op += mlen;
                        __syncwarp();
                    }
                    ip = nip;
                }
            }
            if (blk_cksum) ip += 4;
        }
        if (status == QZB_ST_OK && m.exact_out && op != cap) status = QZB_ST_SIZE;
        uint32_t ck = 0;
        if (status == QZB_ST_OK) {
            __syncwarp();
            ck = warp_xxh32_global(dst, op, lane);
            if (m.check_cksum && ck != m.expect_cksum) status = QZB_ST_CKSUM;
        }
        if (lane == 0) {
            QzbMemberResult r;
            r.status = status; r.consumed = ip; r.produced = op; r.cksum = ck; r.saw_final = 1; r.safe_consumed = r.safe_produced = r.pad = 0;
            job.results[mi] = r;
        }
        __syncwarp();
    }
}

#ifndef QZ_WARP_EMU
extern "C" size_t qzb_lz4_smem_bytes(int piece_log2, int warps)
{
    return (piece_log2 == 13 ? sizeof(Lz4WarpSmem<13>) : sizeof(Lz4WarpSmem<14>)) * (size_t)warps;
}
template <int P>
static cudaError_t launch_lz4(const QzbCompressJob &job, int grid, int warps, cudaStream_t st)
{
    size_t smem = sizeof(Lz4WarpSmem<P>) * (size_t)warps;
    cudaError_t e = cudaFuncSetAttribute(qzb_lz4_pieces_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    qzb_lz4_pieces_kernel<P><<<grid, warps * 32, smem, st>>>(job);
    return cudaGetLastError();
}
/* window kernel with `nw` (12 or 16) warps: the most entries a table can have, shared memory for them, slot scratch for a grid */
extern "C" size_t qzb_lz4_window_smem_bytes(int tent, int nw) { return lz4_window_smem((uint32_t)tent, (uint32_t)nw); }
extern "C" int qzb_lz4_window_max_tent(int nw)
{
    cudaFuncAttributes a;
    if ((nw == 12 ? cudaFuncGetAttributes(&a, qzb_lz4_window_kernel<12>) : cudaFuncGetAttributes(&a, qzb_lz4_window_kernel<16>)) != cudaSuccess) { (void)cudaGetLastError(); return 256; }
    const size_t cap = 227 * 1024 - a.sharedSizeBytes - 64;
    int tent = 256;
    while (tent + 8 <= 32760 && lz4_window_smem((uint32_t)tent + 8, (uint32_t)nw) <= cap) tent += 8;
    return tent;
}
extern "C" size_t qzb_lz4_window_tok_words(int grid, int nw) { return (size_t)grid * nw * QZB_TOK_STRIDE(lz4_sub_bytes((uint32_t)nw)); }
template <int NW>
static cudaError_t launch_lz4_window(const QzbCompressJob &job, int grid, cudaStream_t st)
{
    const size_t smem = lz4_window_smem(job.tent, NW);
    cudaError_t e = cudaFuncSetAttribute(qzb_lz4_window_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    qzb_lz4_window_kernel<NW><<<grid, NW * 32, smem, st>>>(job);
    return cudaGetLastError();
}
extern "C" cudaError_t qzb_launch_lz4_window(const QzbCompressJob *job, int grid, int nw, cudaStream_t st)
{
    if ((nw != 12 && nw != 16) || job->pieces_per_chunk % 8 || !job->ngroups || job->piece_log2 != 13 || job->tent < 256 || job->tent > 32760) return cudaErrorInvalidValue;
    cudaError_t e = nw == 12 ? launch_lz4_window<12>(*job, grid, st) : launch_lz4_window<16>(*job, grid, st);
    if (e != cudaSuccess) return e;
    qzb_xxh32_chunks_kernel<<<(job->nchunks * 4 + 255) / 256, 256, 0, st>>>(*job);
    return cudaGetLastError();
}
extern "C" cudaError_t qzb_launch_lz4_compress(const QzbCompressJob *job, int grid, int warps, cudaStream_t st)
{
    cudaError_t e = job->piece_log2 == 13 ? launch_lz4<13>(*job, grid, warps, st) : launch_lz4<14>(*job, grid, warps, st);
    if (e != cudaSuccess) return e;
    qzb_xxh32_chunks_kernel<<<(job->nchunks * 4 + 255) / 256, 256, 0, st>>>(*job);
    return cudaGetLastError();
}
extern "C" cudaError_t qzb_launch_lz4_decompress(const QzbDecompressJob *job, int grid, cudaStream_t st)
{
    qzb_lz4_decompress_kernel<<<grid, 256, 0, st>>>(*job);
    return cudaGetLastError();
}
#endif
