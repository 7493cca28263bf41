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
            r.status = status; r.consumed = ip; r.produced = op; r.cksum = ck; r.saw_final = 1; r.pad[0] = r.pad[1] = r.pad[2] = 0;
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
