/* qz_deflate.cu -- sm_100a DEFLATE compressor.  The device's unit is the PIECE (8 KiB of a chunk, private window): one
 * warp matches it out of a shared-memory piece buffer with a private u16 hash table, so HBM sees each input byte once
 * and each output byte once (tokens live in an L2-resident scratch in between).
 *
 * These kernels replace the QAT compress request submitted at reference src/qatzip.c:1542 (cpaDcCompressData2, stateless
 * deflate, CPA_DC_FLUSH_FINAL / _FULL) with the session set-up of reference src/qatzip_utils.c:264-341 (dynamic or
 * static Huffman, stored fallback, CRC-32 of the input returned in res.checksum), and the stitching doCompressOut does
 * afterwards (framing kernels at the end of the file).
 *
 * Phases of a piece, all warp-synchronous:
 *   1 load    global -> shared, 16 B per lane, then CRC-32 / Adler-32 over right-aligned per-lane strips
 *   2 match   32 positions per step: 4-byte hash probe, 12-byte verify + extend with all loads in flight, ballot-driven
 *             greedy selection, raw tokens to the scratch                                          (phase12)
 *   3 code    token pass (symbols, histograms); sort by frequency in registers, in-place Huffman lengths (merge on one
 *             lane, depths by pointer doubling, leaf depths across the warp), canonical codes, dynamic header plan;
 *             cheapest of stored / fixed / dynamic                                 (token_pass, choose_block, open_block)
 *   4 emit    every lane packs a contiguous run of tokens at a bit offset from a scan of the runs' lengths; run-boundary
 *             words are zeroed first and joined by atomic OR                          (count_run_bits, emit_run)
 *
 * Two kernels string them together:
 *   qzb_deflate_groups_kernel  (default for hw_buff_sz >= 64 KiB) eight warps = eight pieces = ONE deflate block: phases 1-2
 *                              and the token pass per warp, the code construction once per group by its leader, emission
 *                              at bit offsets inside the group's output; warps of a group meet at a named barrier
 *   qzb_deflate_pieces_kernel  one block (or stored block) per piece, no cross-warp step: smaller chunks, 16 KiB pieces,
 *                              QZB200_GROUP=0
 * A CTA's warps share a pool of piece buffers (held only while matching, or by a group leader as code scratch); the
 * experimental matcher / coder variant lives in qz_deflate_split.cuh.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include "qz_kernels.cuh"
#include "qz_warp.cuh"
#include "qz_huffman.h"
#include "qz_crc32.h"
#include "qz_adler32.h"

#define FULL 0xffffffffu
#define QZ_NONE16 0xffffu
#define QZ_LANE_CAP 12          /* in-lane match extension cap; longer matches are finished by the warp */
#define QZ_MAX_MATCH 258
#define QZ_STAGE_WORDS 64
/* most warps a CTA may have.  Per-piece kernel: shared memory admits 20-24, and the bound lets the compiler use up to
 * 80 registers per thread instead of the 64 a 1024-thread bound would impose.  Group kernel: 2 KiB per warp with the
 * 2^10-entry table, so four groups of eight warps fit beside 17 piece buffers, and that measured faster than 24 warps
 * with 80 registers. */
#ifndef QZ_PIECES_MAX_WARPS
#define QZ_PIECES_MAX_WARPS 24
#endif
#ifndef QZ_GROUPS_MAX_WARPS
#define QZ_GROUPS_MAX_WARPS 32
#endif

/* Per-phase cycle accounting for on-box diagnosis (A/B build only: make ab ABFLAGS=-DQZ_PHASE_CLOCKS).
 * Lane 0 of every warp adds the cycles since its previous mark to a global counter per phase. */
#ifdef QZ_PHASE_CLOCKS
__device__ unsigned long long qz_phase_cycles[16];
#define QZ_MARK(i) do { if (lane == 0) { long long now_ = clock64(); atomicAdd(&qz_phase_cycles[i], (unsigned long long)(now_ - tlast)); tlast = now_; } } while (0)
#define QZ_TARG , long long &tlast
#define QZ_TPASS , tlast
extern "C" __attribute__((visibility("default"))) int qzb_phase_cycles_read(unsigned long long *out, int reset)
{
    if (cudaMemcpyFromSymbol(out, qz_phase_cycles, sizeof(qz_phase_cycles)) != cudaSuccess) return -1;
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(qz_phase_cycles, z, sizeof(z)); }
    return 0;
}
#else
#define QZ_MARK(i) do { } while (0)
#define QZ_TARG
#define QZ_TPASS
#endif

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ uint32_t lanemask_lt() { return qz_lanemask_lt(); }

/* unaligned 32-bit read from a 4-byte aligned shared byte array */
__device__ __forceinline__ uint32_t ld32u(const uint8_t *base, uint32_t off)
{
    const uint32_t *w = reinterpret_cast<const uint32_t *>(base) + (off >> 2);
    return __funnelshift_r(w[0], w[1], (off & 3) * 8);
}

/* Warp bitonic sort of 32*K keys held K per lane (element e = lane * K + r): exchanges at distance
 * >= K are shuffles, shorter ones stay inside the lane's registers. */
template <int K>
__device__ __forceinline__ void warp_sort_regs(uint32_t (&x)[K], uint32_t lane)
{
#pragma unroll 1
    for (int k = 2; k <= 32 * K; k <<= 1) {
#pragma unroll 1
        for (int j = k >> 1; j >= K; j >>= 1) {
            const int m = j / K;
            const bool lower = (lane & m) == 0;
            const bool up = ((lane * K) & k) == 0;        /* j >= K => k > K: the direction bit lies in the lane part */
#pragma unroll
            for (int r = 0; r < K; r++) {
                const uint32_t o = __shfl_xor_sync(FULL, x[r], m);
                x[r] = (lower == up) ? min(x[r], o) : max(x[r], o);
            }
        }
#pragma unroll
        for (int j = K >> 1; j > 0; j >>= 1) {
            if (j < k) {
#pragma unroll
                for (int r = 0; r < K; r++) {
                    if ((r & j) == 0) {
                        const bool up = (((lane * K + r) & k) == 0);
                        const uint32_t a = x[r], c = x[r | j];
                        const uint32_t mn = min(a, c), mx = max(a, c);
                        x[r] = up ? mn : mx; x[r | j] = up ? mx : mn;
                    }
                }
            }
        }
    }
}
/* keys[0..n) (QZ_HUFF_KEY) -> ascending; written back split: freq[e] = frequency, ids[e] = symbol */
template <int K>
__device__ __noinline__ void warp_sort_split(const uint32_t *keys, int n, uint32_t *freq, uint16_t *ids, uint32_t lane)
{
    uint32_t x[K];
#pragma unroll
    for (int r = 0; r < K; r++) { const int e = r * 32 + (int)lane; x[r] = e < n ? keys[e] : 0xffffffffu; }   /* any input order will do */
    __syncwarp();
    warp_sort_regs<K>(x, lane);
#pragma unroll
    for (int r = 0; r < K; r++) { const int e = (int)lane * K + r; if (e < n) { freq[e] = x[r] >> 9; ids[e] = (uint16_t)(x[r] & 511u); } }
    __syncwarp();
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

/* Pass 3 of the in-place Huffman construction across the warp.  A[0..n-1) holds the depths of the
 * internal nodes, non-increasing with the index.  Lane d counts the internal nodes at depth d by two
 * binary searches' worth of work (one, plus its neighbour's by shuffle); a level with I internal nodes
 * offers 2 I places to the next one, and what the next level's internal nodes leave over are leaves.
 * Leaves are handed out from the top of A (most frequent symbol, smallest depth) downwards.
 * Returns false when the tree is deeper than 30 levels (caller falls back to the serial pass). */
__device__ __forceinline__ bool warp_leaf_depths(uint32_t *A, int n, uint32_t lane)
{
    const int ni = n - 1;
    /* c(d) = number of internal nodes with depth >= d = first index whose depth is < d */
    int lo = 0, hi = ni;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (A[mid] >= lane) lo = mid + 1; else hi = mid; }
    const int c_ge = lo;
    const int c_next = __shfl_down_sync(FULL, c_ge, 1);
    if (__shfl_sync(FULL, c_ge, 31) != 0) return false;
    const int I = c_ge - (lane == 31 ? 0 : c_next);
    const int Iprev = __shfl_up_sync(FULL, I, 1);
    const int Lf = (lane == 0 ? 1 : 2 * Iprev) - I;                  /* leaves at depth = lane */
    int incl = Lf;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(FULL, incl, o); if (lane >= (uint32_t)o) incl += y; }
    __syncwarp();
    uint32_t levels = __ballot_sync(FULL, Lf > 0);
    while (levels) {
        const int d = __ffs(levels) - 1; levels &= levels - 1;
        const int cnt = __shfl_sync(FULL, Lf, d), top = n - (__shfl_sync(FULL, incl, d) - cnt);
        for (int i = lane; i < cnt; i += 32) A[top - 1 - i] = (uint32_t)d;
    }
    __syncwarp();
    return true;
}

/* Pass 2 of the in-place Huffman construction across the warp: A[i], i < n - 2, holds the parent of internal node i (always a
 * higher index), node n - 2 is the root.  Pointer doubling on packed (distance so far << 16 | ancestor): every round each
 * node adds its ancestor's distance and adopts the ancestor's ancestor, reads and writes of a round separated by a warp
 * sync, until every node points at the root -- ceil(log2(depth)) rounds instead of n dependent double loads on one lane.
 * On return A[0..n-1) are the depths (root 0). */
template <int K>
__device__ __forceinline__ void warp_depth_pass(uint32_t *A, int n, uint32_t lane)
{
    const int ni = n - 1, r = n - 2;
    uint32_t x[K];
#pragma unroll
    for (int k = 0; k < K; k++) { const int i = (int)lane + 32 * k; x[k] = i < r ? ((1u << 16) | A[i]) : (uint32_t)r; }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < K; k++) { const int i = (int)lane + 32 * k; if (i < ni) A[i] = x[k]; }
    __syncwarp();
    for (;;) {
        bool more = false;
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int i = (int)lane + 32 * k;
            const uint32_t p = x[k] & 0xffffu;
            if (i < r && p != (uint32_t)r) {
                const uint32_t y = A[p];
                x[k] = (((x[k] >> 16) + (y >> 16)) << 16) | (y & 0xffffu);
                more |= (y & 0xffffu) != (uint32_t)r;
            }
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < K; k++) { const int i = (int)lane + 32 * k; if (i < r) A[i] = x[k]; }
        __syncwarp();
        if (__ballot_sync(FULL, more) == 0) break;
    }
#pragma unroll
    for (int k = 0; k < K; k++) { const int i = (int)lane + 32 * k; if (i < ni) A[i] = x[k] >> 16; }
    __syncwarp();
}

/* sorted frequencies -> code lengths per symbol, for the literal/length and the distance alphabet at once:
 * the merge pass is serial, so lane 0 merges the literal/length tree while lane 1 merges the distance tree
 * (branch-free code: the two lanes stay converged); node depths, leaf depths and the scatter back to symbol
 * order run across the warp */
__device__ __noinline__ void warp_lengths_pair(uint32_t *keys, const uint16_t *ids, int n, uint8_t *ll_len,
                                               uint32_t *dkeys, const uint16_t *dids, int nd, uint8_t *d_len, uint32_t lane)
{
    if (lane < 2) qz_huff_merge_pass(lane ? dkeys : keys, lane ? nd : n);
    __syncwarp();
    warp_depth_pass<9>(keys, n, lane);
    warp_depth_pass<1>(dkeys, nd, lane);
    const bool ok_ll = warp_leaf_depths(keys, n, lane);
    const bool ok_d = warp_leaf_depths(dkeys, nd, lane);
    if (lane < 2) {
        uint32_t *A = lane ? dkeys : keys; const int m = lane ? nd : n;
        if (!(lane ? ok_d : ok_ll)) qz_huff_depths_to_lengths(A, m);
        qz_huff_limit_sorted(A, m, 15);
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) ll_len[ids[i]] = (uint8_t)keys[i];
    if ((int)lane < nd) d_len[dids[lane]] = (uint8_t)dkeys[lane];
    __syncwarp();
}

/* canonical codes for len[0..n) -> out[s] = bit-reversed code | len << 16, in symbol order.
 * Symbols are taken 32 at a time; lanes holding equal lengths find each other with match.any
 * and take consecutive codes.  scratch = 32 words. */
__device__ __noinline__ void warp_assign_codes(const uint8_t *len, int n, uint32_t *out, uint32_t *scratch, uint32_t lane)
{
    uint32_t *cnt = scratch, *next = scratch + 16;
    if (lane < 16) cnt[lane] = 0;
    __syncwarp();
    for (int s = lane; s < n; s += 32) { uint32_t l = len[s]; if (l) atomicAdd(&cnt[l], 1u); }
    __syncwarp();
    if (lane == 0) { uint32_t code = 0; for (int l = 1; l < 16; l++) { next[l] = code; code = (code + cnt[l]) << 1; } next[0] = 0; }
    __syncwarp();
    for (int s0 = 0; s0 < n; s0 += 32) {
        const int s = s0 + (int)lane;
        const uint32_t l = s < n ? len[s] : 0u;
        const uint32_t same = __match_any_sync(FULL, l);
        const uint32_t rank = __popc(same & lanemask_lt());
        const uint32_t base = next[l];
        __syncwarp();
        if (l && rank == 0) next[l] = base + __popc(same);
        if (s < n) out[s] = l ? ((__brev(base + rank) >> (32 - l)) | (l << 16)) : 0u;
        __syncwarp();
    }
}

/* Everything phases 3-4 keep in shared memory lives where the hash table was (dead after phase 2):
 * sort keys -> frequencies -> header-plan counters -> bit staging window, symbol ids, code lengths, the
 * dynamic header, and the histograms that later become the code tables.  4000 B against the 4096 B
 * of a 2^11-entry table, so a warp's private slice is exactly its hash table. */
struct CodeScratch {
    uint32_t keys[288];                 /* 286 sort keys at most; see above for its later lives */
    uint16_t ids[QZ_NUM_LL + 2];
    uint8_t ll_len[288];
    uint8_t d_len[32];
    QzDynHeaderCore hdr;
};
#define QZ_HIST_WORDS (QZ_NUM_LL + 2 + QZ_NUM_D + 2)   /* [0,286) lit/len, [288,318) dist; later the code tables */

/* Shared-memory plan.  The piece buffer (8 KiB) is needed only by phases 1-2; phases 3-4 work
 * from the token scratch and the warp's private tables.  So a CTA owns NB piece buffers and
 * NW > NB warps: a warp draws a ticket, takes any free buffer, runs phases 1-2, hands the buffer
 * back and finishes phases 3-4 without it. */
template <int HB>
struct WarpPriv {
    union {
        uint16_t table[1 << HB];
        struct { CodeScratch cs; uint32_t hist[QZ_HIST_WORDS]; } b;
    } u;
};
template <int PIECE_LOG2>
struct PieceBuf {
    uint8_t bytes[(1 << PIECE_LOG2) + 32];  /* +32: zero pad so unaligned reads past n are defined */
};
#define QZ_DOFF 288

/* what phases 3-4 need to know about the piece phases 1-2 just finished */
struct PieceState {
    uint32_t g, n, ntok, extra_total;
    bool bfinal;
    const uint8_t *src;
};

#ifdef QZ_EMU_STATS
unsigned long long qz_stat[8];      /* tiles, skipped tiles, selection iterations, long matches, long-match steps, tokens, matches */
#define QZ_STAT(i, v) do { if (lane == 0) qz_stat[i] += (v); } while (0)
#else
#define QZ_STAT(i, v) do { } while (0)
#endif
template <int PIECE_LOG2, int HB>
__device__ __forceinline__ void phase12(const QzbCompressJob &job, uint8_t *piece, uint16_t *table, uint32_t *toks,
                                        const uint32_t *s_crc_tab, const uint32_t *s_xstrip, uint32_t g, uint32_t lane, PieceState &ps QZ_TARG)
{
    constexpr int PIECE = 1 << PIECE_LOG2;
    constexpr uint32_t STRIP = PIECE / 32 + 4;   /* bytes per lane; /4 is odd -> conflict-free banks */
    /* ---- which bytes ---- */
    const uint32_t chunk = g / job.pieces_per_chunk, k = g - chunk * job.pieces_per_chunk;
    const uint64_t chunk_off = (uint64_t)chunk * job.chunk_sz;
    const uint64_t rem = job.src_len > chunk_off ? job.src_len - chunk_off : 0;
    const uint32_t chunk_len = rem < job.chunk_sz ? (uint32_t)rem : job.chunk_sz;
    const uint32_t p_off = k << PIECE_LOG2;
    const uint32_t n = chunk_len > p_off ? min((uint32_t)PIECE, chunk_len - p_off) : 0u;
    const bool last_piece = (p_off + n == chunk_len);
    ps.g = g; ps.n = n;
    ps.bfinal = last_piece && (job.fmt != QZB_FMT_RAW || (chunk == job.nchunks - 1 && job.last));
    const uint8_t *src = job.src + chunk_off + p_off;
    ps.src = src;

    /* ---- phase 1: load + CRC ---- */
    {
        uint4 *d4 = reinterpret_cast<uint4 *>(piece);
        if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
            const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
            uint32_t nv = n >> 4;
            const uint64_t pstream = l2_policy_stream();
            for (uint32_t i = lane; i < nv; i += 32) d4[i] = stream_ld16(s4 + i, pstream);
            for (uint32_t i = (nv << 4) + lane; i < n; i += 32) piece[i] = src[i];
        } else {
            for (uint32_t i = lane; i < n; i += 32) piece[i] = src[i];
        }
        piece[n + lane] = 0;          /* zero pad */
        for (uint32_t i = lane; i < (1u << HB) / 2; i += 32) reinterpret_cast<uint32_t *>(table)[i] = 0xffffffffu;
        __syncwarp();
        /* right-aligned strips: lane i owns [n-(32-i)*STRIP, n-(31-i)*STRIP) clipped at 0 */
        int hi = (int)n - (int)((31 - lane) * STRIP), lo = hi - (int)STRIP;
        if (lo < 0) lo = 0;
        uint32_t c;
        if (job.fmt == QZB_FMT_ZLIB) {
            /* Adler-32 sums of the strip (STRIP < NMAX: no reduction inside), joined up the same tree */
            uint32_t s1 = 0, s2 = 0;
            for (int i = lo; i < hi; i++) { s1 += piece[i]; s2 += s1; }
            s2 %= QZ_ADLER_P;
#pragma unroll 1
            for (int lv = 0; lv < 5; lv++) {
                const uint32_t o1 = __shfl_down_sync(FULL, s1, 1u << lv), o2 = __shfl_down_sync(FULL, s2, 1u << lv);
                if ((lane & ((2u << lv) - 1)) == 0) qz_adler_join(&s1, &s2, o1, o2, (uint64_t)STRIP << lv);
            }
            c = qz_adler_pack(s1, s2);
        } else {
            c = 0xffffffffu;
            for (int i = lo; i < hi; i++) c = s_crc_tab[(c ^ piece[i]) & 0xff] ^ (c >> 8);
            c = (hi > lo) ? ~c : 0u;           /* empty strip -> CRC of nothing */
#pragma unroll
            for (int lv = 0; lv < 5; lv++) {
                uint32_t other = __shfl_down_sync(FULL, c, 1u << lv);   /* right neighbour block */
                if ((lane & ((2u << lv) - 1)) == 0) c = qz_gf2_mul(c, s_xstrip[lv]) ^ other;
            }
        }
        if (lane == 0) job.piece_crc[g] = c;
    }
    QZ_MARK(1);

    /* ---- phase 2: match + select + tokens ---- */
    uint32_t ntok = 0;
    const uint64_t pkeep = l2_policy_keep();
    {
        uint32_t entry = 0;                 /* first position of the tile not covered by a previous match */
        const uint32_t sh = (lane & 3) * 8;  /* base is a multiple of 32: the byte phase of p is the lane's */
        const uint32_t *piece_w = reinterpret_cast<const uint32_t *>(piece);
        for (uint32_t base = 0; base < n; base += 32) {
            QZ_STAT(0, 1);
            if (entry >= 32) { entry -= 32; QZ_STAT(1, 1); continue; }      /* tile lies inside a running match: nothing to code, not indexed */
            const uint32_t p = base + lane;
            /* 12 bytes at p, straight-line: the three words every lane needs for verify + in-lane extension */
            const uint32_t *pw = piece_w + (min(p, n) >> 2);
            const uint32_t w0 = pw[0], w1 = pw[1], w2 = pw[2], w3 = pw[3];
            const uint32_t v = __funnelshift_r(w0, w1, sh), v1 = __funnelshift_r(w1, w2, sh), v2 = __funnelshift_r(w2, w3, sh);
            const bool can = p + 4 <= n;
            const uint32_t h = (v * 2654435761u) >> (32 - HB);
            uint32_t cand = can ? table[h] : QZ_NONE16;
            __syncwarp();
            if (can) table[h] = (uint16_t)p;
            __syncwarp();
            /* candidate bytes are fetched unconditionally (slot 0 when there is none): no divergent
             * verify/extend branches, all loads in flight together */
            const bool has = cand != QZ_NONE16;
            const uint32_t c = has ? cand : 0u, csh = (c & 3) * 8;
            const uint32_t *cw = piece_w + (c >> 2);
            const uint32_t c0 = cw[0], c1 = cw[1], c2 = cw[2], c3 = cw[3];
            const uint32_t x0 = __funnelshift_r(c0, c1, csh) ^ v, x1 = __funnelshift_r(c1, c2, csh) ^ v1, x2 = __funnelshift_r(c2, c3, csh) ^ v2;
            uint32_t L = x1 ? 4 + ((__ffs(x1) - 1) >> 3) : x2 ? 8 + ((__ffs(x2) - 1) >> 3) : (uint32_t)QZ_LANE_CAP;
            L = min(L, min((uint32_t)QZ_MAX_MATCH, n - min(p, n)));
            if (!has || x0) L = 0;
            const uint32_t valid = __ballot_sync(FULL, p < n);
            const uint32_t M = __ballot_sync(FULL, L >= 4);
            uint32_t matchmask = 0, cur = entry;
            for (;;) {
                const uint32_t rest = M & (FULL << cur);
                if (!rest) { cur = 32; break; }
                const uint32_t m = __ffs(rest) - 1;
                QZ_STAT(2, 1);
                matchmask |= 1u << m;
                uint32_t Lm = __shfl_sync(FULL, L, m);
                if (Lm >= QZ_LANE_CAP) {                      /* warp finishes the long match */
                    const uint32_t pm = base + m, cm = __shfl_sync(FULL, cand, m);
                    const uint32_t mx = min((uint32_t)QZ_MAX_MATCH, n - pm);
                    Lm = QZ_LANE_CAP;
                    QZ_STAT(3, 1);
                    while (Lm < mx) {
                        QZ_STAT(4, 1);
                        uint32_t kk = Lm + lane;
                        bool eq = kk < mx && piece[pm + kk] == piece[cm + kk];
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
            /* positions covered by the match running in from the previous tile or by a selected match
             * of this one carry no token: every selected lane marks (lane, lane + L), one warp OR joins them */
            const bool sel = (matchmask >> lane) & 1;
            const uint32_t mycov = sel ? (((FULL << lane) << 1) & ~(lane + L >= 32 ? 0u : FULL << (lane + L))) : 0u;
            const uint32_t tokmask = ~(__reduce_or_sync(FULL, mycov) | ~(FULL << entry)) & valid;
            entry = cur - 32;
            if ((tokmask >> lane) & 1) {
                /* raw token: literal byte, or match flag | (len - 3) << 16 | (dist - 1); symbols and
                 * histograms are derived in phase 3, off the piece buffer's critical path */
                const uint32_t t = sel ? (0x80000000u | ((L - 3) << 16) | (p - cand - 1)) : (v & 0xff);
                tok_st(toks + ntok + __popc(tokmask & lanemask_lt()), t, pkeep);
            }
            ntok += __popc(tokmask);
            QZ_STAT(5, __popc(tokmask)); QZ_STAT(6, __popc(matchmask));
        }
    }
    __syncwarp();
    QZ_MARK(2);
    ps.ntok = ntok;
    ps.extra_total = 0;
}

/* ---- bit emission: 32 variable-length fields per call --------------------------------------
 * Each lane contributes `nb` bits (<= 48).  A warp scan gives every field its bit offset, the
 * fields are OR-ed into a zeroed shared staging window, and every word that became complete is
 * stored (coalesced) to the slot.  st[0] always holds the pending partial word. */
struct EmitState { uint32_t bitpos, flushed; };

__device__ __forceinline__ void emit_group(uint32_t *st, uint32_t *slotw, EmitState &es, uint64_t bits, uint32_t nb, uint32_t lane)
{
    uint32_t incl = nb;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(FULL, incl, o); if (lane >= (uint32_t)o) incl += y; }
    const uint32_t total = __shfl_sync(FULL, incl, 31);
    if (nb) {
        const uint32_t pos = es.bitpos + incl - nb - (es.flushed << 5);
        const uint32_t w = pos >> 5, sh = pos & 31;
        const uint64_t lo = bits << sh;
        const uint32_t hi = sh ? (uint32_t)(bits >> (64 - sh)) : 0u;
        if ((uint32_t)lo) atomicOr(&st[w], (uint32_t)lo);
        if ((uint32_t)(lo >> 32)) atomicOr(&st[w + 1], (uint32_t)(lo >> 32));
        if (hi) atomicOr(&st[w + 2], hi);
    }
    __syncwarp();
    es.bitpos += total;
    const uint32_t nfull = (es.bitpos >> 5) - es.flushed;
    for (uint32_t i = lane; i < nfull; i += 32) slotw[es.flushed + i] = st[i];
    const uint32_t carry = st[nfull];
    __syncwarp();
    for (uint32_t i = lane; i <= nfull + 2 && i < QZ_STAGE_WORDS; i += 32) st[i] = 0;
    __syncwarp();
    if (lane == 0) st[0] = carry;
    __syncwarp();
    es.flushed += nfull;
}

/* length 3..258 -> symbol index 0..28 | extra-bit count << 5 | (base - 3) << 8, filled once per CTA */
__device__ __forceinline__ uint16_t len_table_entry(uint32_t l /* len - 3 */)
{
    uint32_t sym, eb, ev;
    qz_len_code(l + 3, &sym, &eb, &ev);
    return (uint16_t)((sym - 257) | (eb << 5) | ((l - ev) << 8));
}

/* Warp-parallel version of qz_dyn_header_plan's run-length pass (RFC 1951 3.2.7): every run of
 * equal code lengths is handled by the lane sitting on its first element. */
__device__ __noinline__ void warp_plan_header(CodeScratch &cs, uint32_t *cf /* 19 counters + 12 words + 80 words of run-length scratch */, uint32_t lane)
{
    QzDynHeaderCore &h = cs.hdr;
    uint32_t bal = __ballot_sync(FULL, lane < 29 && cs.ll_len[257 + lane] != 0);
    const uint32_t hlit = 257 + (bal ? 32 - __clz(bal) : 0);
    bal = __ballot_sync(FULL, lane < QZ_NUM_D && cs.d_len[lane] != 0);
    const uint32_t hdist = bal ? 32 - __clz(bal) : 1;
    const uint32_t total = hlit + hdist, ngroups = (total + 31) >> 5;
    uint8_t *seq = reinterpret_cast<uint8_t *>(cf + 32);   /* 316 bytes of the dead sort-key space */
    uint32_t *heads = cf + 20;                    /* one head mask per group of 32 positions */
    if (lane < QZ_NUM_CL) cf[lane] = 0;
    for (uint32_t i = lane; i < total; i += 32) seq[i] = i < hlit ? cs.ll_len[i] : cs.d_len[i - hlit];
    __syncwarp();
    for (uint32_t g = 0; g < ngroups; g++) {
        const uint32_t i = (g << 5) + lane;
        const bool head = i < total && (i == 0 || seq[i] != seq[i - 1]);
        const uint32_t m = __ballot_sync(FULL, head);
        if (lane == 0) heads[g] = m;
    }
    __syncwarp();
    uint32_t nitems = 0;
    for (uint32_t g = 0; g < ngroups; g++) {
        const uint32_t i = (g << 5) + lane, hm = heads[g];
        const bool head = (hm >> lane) & 1;
        uint32_t cnt = 0, v = 0, run = 0;
        if (head) {
            v = seq[i];
            uint32_t nxt = total, m = lane < 31 ? hm & (FULL << (lane + 1)) : 0u;
            if (m) nxt = (g << 5) + __ffs(m) - 1;
            else for (uint32_t g2 = g + 1; g2 < ngroups; g2++) { uint32_t m2 = heads[g2]; if (m2) { nxt = (g2 << 5) + __ffs(m2) - 1; break; } }
            run = nxt - i;
            if (v == 0) {
                const uint32_t full = run / 138, rem = run - full * 138;
                cnt = full + (rem >= 11 ? 1 : rem >= 3 ? 1 : rem);
            } else {
                const uint32_t r1 = run - 1, full = r1 / 6, rem = r1 - full * 6;
                cnt = 1 + full + (rem >= 3 ? 1 : rem);
            }
        }
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(FULL, incl, o); if (lane >= (uint32_t)o) incl += y; }
        if (head) {
            uint16_t *it = h.items + nitems + incl - cnt;
            if (v == 0) {
                uint32_t r = run;
                while (r >= 11) { const uint32_t t = r > 138 ? 138 : r; *it++ = qz_cl_item(18, t - 11, 7); atomicAdd(&cf[18], 1u); r -= t; }
                if (r >= 3) { *it++ = qz_cl_item(17, r - 3, 3); atomicAdd(&cf[17], 1u); r = 0; }
                if (r) { atomicAdd(&cf[0], r); while (r--) *it++ = qz_cl_item(0, 0, 0); }
            } else {
                uint32_t r = run - 1, lits = 1;
                *it++ = qz_cl_item(v, 0, 0);
                while (r >= 3) { const uint32_t t = r > 6 ? 6 : r; *it++ = qz_cl_item(16, t - 3, 2); atomicAdd(&cf[16], 1u); r -= t; }
                lits += r;
                while (r--) *it++ = qz_cl_item(v, 0, 0);
                atomicAdd(&cf[v], lits);
            }
        }
        nitems += __shfl_sync(FULL, incl, 31);
    }
    __syncwarp();
    if (lane == 0) { h.hlit = hlit; h.hdist = hdist; h.nitems = nitems; qz_cl_build(cf, &h, cf + 32 /* the run-length scratch is dead by now */); }
    __syncwarp();
}

/* ---- phases 3-4 as building blocks (the per-piece kernel strings them together for one piece; the group
 * kernel runs the token pass and the emission per piece and the code construction once per group) ---- */

/* phase 3a: token pass -- symbols, histograms, and tokens rewritten in symbol form:
 *      match flag | len symbol (5) << 26 | len extra (5) << 21 | dist symbol (5) << 16 | dist extra (13)
 * Returns the piece's total of extra bits. */
__device__ __forceinline__ uint32_t token_pass(uint32_t *hist, uint32_t *toks, uint32_t ntok, const uint16_t *s_lentab, uint32_t lane, uint64_t pkeep, bool zero = true)
{
    if (zero) { for (uint32_t i = lane; i < QZ_HIST_WORDS; i += 32) hist[i] = 0; }
    __syncwarp();
    uint32_t extra_acc = 0;
    uint32_t tnext = lane < ntok ? tok_ld(toks + lane, pkeep) : 0u;          /* one group ahead: hides the L2 round trip */
    for (uint32_t t0 = 0; t0 < ntok; t0 += 32) {
        const uint32_t t = tnext;
        if (t0 + 32 + lane < ntok) tnext = tok_ld(toks + t0 + 32 + lane, pkeep);
        if (t0 + lane < ntok) {
            if (t & 0x80000000u) {
                const uint32_t le = s_lentab[(t >> 16) & 0xff];
                const uint32_t ls = le & 31, leb = (le >> 5) & 7, lev = ((t >> 16) & 0xff) - (le >> 8);
                uint32_t ds, de, dv;
                qz_dist_code((t & 0xffff) + 1, &ds, &de, &dv);
                atomicAdd(&hist[257 + ls], 1u);
                atomicAdd(&hist[QZ_DOFF + ds], 1u);
                extra_acc += leb + de;
                tok_st(toks + t0 + lane, 0x80000000u | (ls << 26) | (lev << 21) | (ds << 16) | dv, pkeep);
            } else atomicAdd(&hist[t], 1u);
        }
    }
    return extra_acc;
}

/* phase 3b: code construction from the histograms in hist -> code lengths in cs.ll_len / cs.d_len, the planned
 * dynamic header in cs.hdr, and the cheapest block type for `storedb` bits of stored cost: 0 stored, 1 fixed, 2 dynamic */
__device__ __forceinline__ int choose_block(CodeScratch &cs, uint32_t *hist, uint32_t extra_total, uint32_t storedb, int static_huffman, uint32_t lane QZ_TARG)
{
    for (uint32_t i = lane; i < 288; i += 32) cs.ll_len[i] = 0;
    cs.d_len[lane] = 0;
    if (lane == 0) {
        hist[256] = 1;
        qz_huff_force_two(hist, QZ_NUM_LL);
        qz_huff_force_two(hist + QZ_DOFF, QZ_NUM_D);
    }
    __syncwarp();
    /* literal/length alphabet */
    int nk = 0;
    for (uint32_t s0 = 0; s0 < 288; s0 += 32) {
        uint32_t s = s0 + lane, f = s < QZ_NUM_LL ? hist[s] : 0;
        uint32_t bal = __ballot_sync(FULL, f != 0);
        if (f) cs.keys[nk + __popc(bal & lanemask_lt())] = QZ_HUFF_KEY(f, s);
        nk += __popc(bal);
    }
    __syncwarp();
    /* sort in registers; frequencies land back in keys[], symbols in ids[] */
    if (nk <= 128) warp_sort_split<4>(cs.keys, nk, cs.keys, cs.ids, lane);
    else if (nk <= 256) warp_sort_split<8>(cs.keys, nk, cs.keys, cs.ids, lane);
    else warp_sort_split<16>(cs.keys, nk, cs.keys, cs.ids, lane);
    /* distance alphabet: one key per lane; its frequencies and ids borrow the (not yet planned) header area */
    {
        uint32_t *dkeys = reinterpret_cast<uint32_t *>(cs.hdr.items);
        uint16_t *dids = cs.hdr.items + 64;
        const uint32_t f = lane < QZ_NUM_D ? hist[QZ_DOFF + lane] : 0;
        const int nd = __popc(__ballot_sync(FULL, f != 0));
        uint32_t x[1] = { f ? QZ_HUFF_KEY(f, lane) : 0xffffffffu };
        warp_sort_regs<1>(x, lane);
        if ((int)lane < nd) { dkeys[lane] = x[0] >> 9; dids[lane] = (uint16_t)(x[0] & 511u); }
        __syncwarp();
        QZ_MARK(4);
        warp_lengths_pair(cs.keys, cs.ids, nk, cs.ll_len, dkeys, dids, nd, cs.d_len, lane);
        QZ_MARK(5);
    }
    /* cost of each block type */
    uint32_t dynb = 0, fixb = 0;
    for (uint32_t s = lane; s < QZ_NUM_LL; s += 32) { uint32_t f = hist[s]; dynb += f * cs.ll_len[s]; fixb += f * qz_fixed_ll_len(s); }
    if (lane < QZ_NUM_D) { uint32_t f = hist[QZ_DOFF + lane]; dynb += f * cs.d_len[lane]; fixb += f * 5; }
    dynb = warp_sum(dynb) + extra_total; fixb = warp_sum(fixb) + extra_total + 3;
    /* forced dummy symbols were counted with freq 1 but are never emitted: harmless overestimate */
    warp_plan_header(cs, cs.keys, lane);
    QZ_MARK(6);
    dynb += cs.hdr.bits;
    if (static_huffman) dynb = 0xffffffffu;
    return (dynb <= fixb && dynb < storedb) ? 2 : (fixb < storedb ? 1 : 0);
}

/* phase 3c: open a fixed (btype 1) or dynamic (2) block at the start of slotw: code tables go where the histograms
 * were (code | len << 16 | extra-bit count << 24), the block header is written; *hb = bits written so far, *pend = the
 * partial word at that position (the first token run continues it) */
__device__ __forceinline__ void open_block(CodeScratch &cs, uint32_t *hist, int btype, bool bfinal, uint32_t *slotw, uint32_t lane, uint32_t *hb, uint32_t *pend_out QZ_TARG)
{
    QzBitWriter bw; bw.acc = 0;
    EmitState es; es.bitpos = 0; es.flushed = 0;
    if (btype == 1) {
        for (uint32_t s = lane; s < 288; s += 32) cs.ll_len[s] = (uint8_t)qz_fixed_ll_len(s);
        cs.d_len[lane] = 5;
        __syncwarp();
    }
    warp_assign_codes(cs.ll_len, 288, hist, cs.keys, lane);
    warp_assign_codes(cs.d_len, btype == 1 ? 32 : QZ_NUM_D, hist + QZ_DOFF, cs.keys, lane);
    uint32_t *clc = cs.keys + 40;                 /* code-length alphabet codes, 19 words */
    if (btype == 2) warp_assign_codes(cs.hdr.cl_len, QZ_NUM_CL, clc, cs.keys, lane);
    if (lane == 0) {
        qz_bw_init(&bw, slotw);
        if (btype == 2) qz_dyn_header_write_prefix(&bw, &cs.hdr, bfinal);
        else qz_bw_put(&bw, (bfinal ? 1u : 0u) | (1u << 1), 3);
        es.bitpos = qz_bw_bitpos(&bw); es.flushed = bw.wpos;
    }
    es.bitpos = __shfl_sync(FULL, es.bitpos, 0); es.flushed = __shfl_sync(FULL, es.flushed, 0);
    const uint32_t pend = __shfl_sync(FULL, (uint32_t)bw.acc, 0);
    /* the cl codes must survive the staging window being cleared: keep them in registers */
    const uint32_t my_clc = (btype == 2 && lane < QZ_NUM_CL) ? clc[lane] : 0u;
    __syncwarp();
    uint32_t *st = cs.keys;                      /* staging window, QZ_STAGE_WORDS words */
    for (uint32_t i = lane; i < QZ_STAGE_WORDS; i += 32) st[i] = 0;
    __syncwarp();
    if (lane == 0) st[0] = pend;
    __syncwarp();

    QZ_MARK(7);
    /* run-length coded code lengths of a dynamic header go through the same emitter */
    if (btype == 2) {
        const uint32_t nitems = cs.hdr.nitems;
        for (uint32_t k0 = 0; k0 < nitems; k0 += 32) {
            uint64_t bits = 0; uint32_t nb = 0;
            const uint32_t it = k0 + lane < nitems ? cs.hdr.items[k0 + lane] : 0xffffffffu;
            const uint32_t c = __shfl_sync(FULL, my_clc, it & 31);
            if (it != 0xffffffffu) { bits = c & 0xffff; nb = c >> 16; bits |= (uint64_t)((it >> 5) & 127) << nb; nb += (it >> 12) & 15; }
            emit_group(st, slotw, es, bits, nb, lane);
        }
    }
    __syncwarp();
    *pend_out = st[0]; *hb = es.bitpos;
    /* code table entries gain their extra-bit counts: code | len << 16 | extra << 24 */
    if (lane < 29) hist[257 + lane] |= ((lane < 8 || lane == 28) ? 0u : (lane - 4) >> 2) << 24;
    if (lane < QZ_NUM_D) hist[QZ_DOFF + lane] |= (lane < 4 ? 0u : (lane >> 1) - 1) << 24;
    if (lane >= QZ_NUM_D) hist[QZ_DOFF + lane] = 0;            /* entry 31: "no distance part" for literals */
    __syncwarp();
}

/* phase 4, pass 1: bits the tokens [beg, end) take under the code tables `tab` */
__device__ __forceinline__ uint32_t count_run_bits(const uint32_t *tab, const uint32_t *toks, uint32_t beg, uint32_t end, uint64_t pkeep)
{
    uint32_t mybits = 0;
    uint4 qn = beg < end ? tok_ld4(toks + beg, pkeep) : make_uint4(0, 0, 0, 0);   /* one group ahead: hides the L2 round trip */
    for (uint32_t j = beg; j < end; j += 4) {
        const uint32_t tt[4] = { qn.x, qn.y, qn.z, qn.w };
        if (j + 4 < end) qn = tok_ld4(toks + j + 4, pkeep);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (j + k < end) {
                const uint32_t t = tt[k];
                const bool isM = (t >> 31) != 0;
                const uint32_t c1 = tab[isM ? 257 + ((t >> 26) & 31) : (t & 511)];
                const uint32_t c2 = tab[QZ_DOFF + (isM ? (t >> 16) & 31 : 31u)];
                mybits += ((c1 >> 16) & 0xff) + (c1 >> 24) + ((c2 >> 16) & 0xff) + (c2 >> 24);
            }
        }
    }
    return mybits;
}

/* phase 4, pass 2: pack the tokens [beg, end) at bit `start` of slotw through a private 64-bit accumulator.  `acc0` is
 * what already sits in the first word below the start bit and `owns_first` says this lane writes that word whole (it
 * continues the block header); every other lane joins its first word by atomic OR when it starts inside one.  With
 * `trailer` the lane appends the byte-aligning empty stored block (nz zero bits, 00 00 ff ff) after its last token. */
__device__ __forceinline__ void emit_run(const uint32_t *tab, const uint32_t *toks, uint32_t beg, uint32_t end, uint32_t start, uint32_t acc0,
                                         bool owns_first, bool trailer, uint32_t nz, uint32_t *slotw, uint64_t pkeep)
{
    uint64_t acc = owns_first ? (uint64_t)acc0 : 0ull;
    uint32_t nacc = start & 31, wpos = start >> 5;
    bool partial = !owns_first && nacc != 0;
#define QZ_EMIT_FLUSH() do { if (nacc >= 32) { if (partial) { atomicOr(slotw + wpos, (uint32_t)acc); partial = false; } else slot_st(slotw + wpos, (uint32_t)acc); \
                                               acc >>= 32; nacc -= 32; wpos++; } } while (0)
    uint4 qn = beg < end ? tok_ld4(toks + beg, pkeep) : make_uint4(0, 0, 0, 0);
    for (uint32_t j = beg; j < end; j += 4) {
        const uint32_t tt[4] = { qn.x, qn.y, qn.z, qn.w };
        if (j + 4 < end) qn = tok_ld4(toks + j + 4, pkeep);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (j + k < end) {
                const uint32_t t = tt[k];
                const bool isM = (t >> 31) != 0;
                const uint32_t c1 = tab[isM ? 257 + ((t >> 26) & 31) : (t & 511)];
                const uint32_t c2 = tab[QZ_DOFF + (isM ? (t >> 16) & 31 : 31u)];
                const uint32_t l1 = (c1 >> 16) & 0xff, l2 = (c2 >> 16) & 0xff;
                const uint32_t lv = isM ? (t >> 21) & 31 : 0u, dv = isM ? t & 0x1fff : 0u;
                acc |= (uint64_t)((c1 & 0xffff) | (lv << l1)) << nacc; nacc += l1 + (c1 >> 24);
                QZ_EMIT_FLUSH();
                acc |= (uint64_t)((c2 & 0xffff) | (dv << l2)) << nacc; nacc += l2 + (c2 >> 24);
                QZ_EMIT_FLUSH();
            }
        }
    }
    if (trailer) {          /* owner of the end-of-block token: empty stored block */
        nacc += nz; QZ_EMIT_FLUSH();
        nacc += 16; QZ_EMIT_FLUSH();
        acc |= (uint64_t)0xffffu << nacc; nacc += 16; QZ_EMIT_FLUSH();
    }
    if ((uint32_t)acc) atomicOr(slotw + wpos, (uint32_t)acc);
#undef QZ_EMIT_FLUSH
}

/* stored block for one piece: the piece starts byte-aligned, so the 3 header bits + pad are one byte.
 * The bytes come from global memory again (the shared piece buffer already belongs to another warp);
 * incompressible pieces are the only ones that pay this second read. */
__device__ __forceinline__ uint32_t stored_piece(uint8_t *slot, const uint8_t *src, uint32_t n, bool bfinal, uint32_t lane)
{
    if (lane == 0) { slot[0] = bfinal ? 1 : 0; slot[1] = (uint8_t)n; slot[2] = (uint8_t)(n >> 8); slot[3] = (uint8_t)~n; slot[4] = (uint8_t)(~n >> 8); }
    for (uint32_t i0 = lane; i0 < n; i0 += 256) {           /* eight loads in flight per lane */
        uint8_t v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = i0 + 32 * k < n ? src[i0 + 32 * k] : (uint8_t)0;
#pragma unroll
        for (int k = 0; k < 8; k++) if (i0 + 32 * k < n) slot[5 + i0 + 32 * k] = v[k];
    }
    return 5 + n;
}

/* one piece as its own block (or run of blocks): everything after the token pass */
__device__ __forceinline__ void finish_piece(const QzbCompressJob &job, CodeScratch &cs, uint32_t *hist, uint32_t *toks, uint32_t lane, const PieceState &ps,
                                             uint32_t extra_total, uint64_t pkeep QZ_TARG)
{
    const uint32_t g = ps.g, n = ps.n, ntok = ps.ntok;
    const bool bfinal = ps.bfinal;
    uint8_t *slot = job.slots + (size_t)g * job.slot_stride;
    uint32_t *slotw = reinterpret_cast<uint32_t *>(slot);

    uint32_t out_bytes = 0;
    int btype = choose_block(cs, hist, extra_total, (5 + n) * 8, job.static_huffman, lane QZ_TPASS);
    if (n == 0) btype = 1;

    if (btype == 0) out_bytes = stored_piece(slot, ps.src, n, bfinal, lane);
    else {
        uint32_t hb, pend2;
        open_block(cs, hist, btype, bfinal, slotw, lane, &hb, &pend2 QZ_TPASS);
        /* ---- phase 4: emit ----
         * Every lane codes a contiguous run of tokens: pass 1 adds up the run's bit length, a warp scan
         * turns the lengths into bit offsets, pass 2 packs the run through a private 64-bit accumulator
         * straight into the slot.  Words that hold a run boundary are zeroed first and receive their
         * parts by atomic OR; every other word is written whole by exactly one lane.  The end-of-block
         * code is the last token; the lane that owns it also appends the byte-alignment trailer. */
        const uint32_t NT = ntok + 1;
        const uint32_t R = (((NT + 31) >> 5) + 3) & ~3u;
        const uint32_t beg = min(lane * R, NT), end = min(beg + R, NT);
        const uint32_t mybits = count_run_bits(hist, toks, beg, end, pkeep);
        uint32_t incl = mybits;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(FULL, incl, o); if (lane >= (uint32_t)o) incl += y; }
        const uint32_t start = hb + incl - mybits;
        const uint32_t end_bit = hb + __shfl_sync(FULL, incl, 31);
        const uint32_t nz = 3 + ((0u - (end_bit + 3)) & 7);                /* stored-block header + pad to a byte */
        const uint32_t end_bit2 = bfinal ? end_bit : end_bit + nz + 32;
        slotw[start >> 5] = 0;
        if (lane == 31) slotw[end_bit2 >> 5] = 0;
        __syncwarp();
        emit_run(hist, toks, beg, end, start, pend2, lane == 0, !bfinal && beg < NT && end == NT, nz, slotw, pkeep);
        out_bytes = (end_bit2 + 7) >> 3;
        __syncwarp();
    }
    if (lane == 0) job.piece_len[g] = out_bytes;
    __syncwarp();
    QZ_MARK(8);
}

template <int HB>
__device__ __forceinline__ void phase34(const QzbCompressJob &job, WarpPriv<HB> &ws, uint32_t *toks, const uint16_t *s_lentab,
                                        uint32_t lane, const PieceState &ps QZ_TARG)
{
    const uint64_t pkeep = l2_policy_keep();
    const uint32_t extra_acc = token_pass(ws.u.b.hist, toks, ps.ntok, s_lentab, lane, pkeep);
    if (lane == 0) tok_st(toks + ps.ntok, 256u, pkeep);      /* end-of-block rides along as the last token */
    __syncwarp();
    const uint32_t extra_total = warp_sum(extra_acc);
    QZ_MARK(3);
    finish_piece(job, ws.u.b.cs, ws.u.b.hist, toks, lane, ps, extra_total, pkeep QZ_TPASS);
}

template <int PIECE_LOG2, int HB>
__global__ void __launch_bounds__(QZ_PIECES_MAX_WARPS * 32) qzb_deflate_pieces_kernel(QzbCompressJob job, int nbuf)
{
    constexpr int PIECE = 1 << PIECE_LOG2;
    static_assert(sizeof(WarpPriv<HB>) == (sizeof(uint16_t) << HB), "phase 3-4 scratch must fit in the hash table");
    QZ_DYN_SMEM(smem_raw);
    __shared__ uint32_t s_crc_tab[256];
    __shared__ uint32_t s_xstrip[5];        /* x^(8*STRIP*2^k) for the CRC tree */
    __shared__ uint16_t s_lentab[256];
    __shared__ uint32_t s_busy[1];          /* free mask of the piece buffers */
    constexpr uint32_t STRIP = PIECE / 32 + 4;

    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    PieceBuf<PIECE_LOG2> *bufs = reinterpret_cast<PieceBuf<PIECE_LOG2> *>(smem_raw);
    WarpPriv<HB> &ws = reinterpret_cast<WarpPriv<HB> *>(smem_raw + (size_t)nbuf * sizeof(PieceBuf<PIECE_LOG2>))[warp];
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) { s_crc_tab[i] = qz_crc_table_entry(i); s_lentab[i] = len_table_entry(i); }
    if (threadIdx.x < 5) s_xstrip[threadIdx.x] = qz_crc_xpow8((uint64_t)STRIP << threadIdx.x);
    if (threadIdx.x == 0) s_busy[0] = nbuf >= 32 ? FULL : (1u << nbuf) - 1;   /* bit b set = piece buffer b is free */
    __syncthreads();

    const uint32_t gwarp = blockIdx.x * nwarps + warp;
    uint32_t *toks = job.tok_scratch + (size_t)gwarp * QZB_TOK_STRIDE(PIECE);

#ifdef QZ_PHASE_CLOCKS
    long long tlast = clock64();
#endif
    for (;;) {
        uint32_t g = 0;
        if (lane == 0) g = atomicAdd(job.ticket, 1u);
        g = __shfl_sync(FULL, g, 0);
        if (g >= job.npieces) break;
        /* take a free piece buffer: one bit per buffer in s_free; a warp that finds none sleeps with
         * exponential back-off instead of spinning on the issue slots the working warps need */
        uint32_t b = 0;
        if (lane == 0) {
            uint32_t ns = 128;
            for (;;) {
                const uint32_t m = *reinterpret_cast<volatile uint32_t *>(&s_busy[0]);
                if (m) {
                    b = __ffs(m) - 1;
                    if (atomicAnd(&s_busy[0], ~(1u << b)) & (1u << b)) break;
                    continue;
                }
                __nanosleep(ns);
                if (ns < 4096) ns <<= 1;
            }
            __threadfence_block();
        }
        b = __shfl_sync(FULL, b, 0);
        QZ_MARK(0);
        PieceState ps;
        phase12<PIECE_LOG2, HB>(job, bufs[b].bytes, ws.u.table, toks, s_crc_tab, s_xstrip, g, lane, ps QZ_TPASS);
        __syncwarp();
        if (lane == 0) { __threadfence_block(); atomicOr(&s_busy[0], 1u << b); }
        phase34<HB>(job, ws, toks, s_lentab, lane, ps QZ_TPASS);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Group kernel: QZ_GROUP consecutive pieces of one chunk (64 KiB with 8 KiB pieces) become ONE deflate block.
 * Eight warps take a group together.  Each runs phases 1-2 and the token pass on its own piece exactly as above
 * (private window, private token scratch), then the group's leader adds up the eight histograms and does the
 * serial work once -- sort, Huffman lengths, header plan, block type, canonical codes, block header -- and every
 * warp packs its tokens with the shared code tables at its bit offset inside the group's output, which starts at the
 * first piece's slot.  Against one block per piece this divides the per-block work (a quarter of a piece's time)
 * and the per-block bytes (dynamic header, flush marker) by eight; the stream is what a zlib deflate with a
 * Z_FULL_FLUSH per 64 KiB would look like.  Warps of a group meet at a named barrier (bar.sync id, 256); groups of
 * one CTA are independent of each other.  Used when a chunk is a whole number of groups (hw_buff_sz >= 64 KiB). */
#define QZ_GROUP 8                      /* pieces per block */
/* Shared memory of the group kernel: a warp's private slice is its hash table during the match phase and its
 * histogram afterwards (no code scratch: only the leader builds codes), so a 2^10-entry table makes it 2 KiB.  The
 * code scratch and the group's histogram / code tables (GroupLead, 4000 B) live in a piece buffer that the leader
 * takes from the pool for the time the block is being coded: no shared memory is set aside for them. */
template <int HB>
struct GroupWarpPriv {
    union {
        uint16_t table[1 << HB];
        uint32_t hist[QZ_HIST_WORDS];
        uint8_t pad[(((2 << HB) > QZ_HIST_WORDS * 4 ? (2 << HB) : QZ_HIST_WORDS * 4) + 15) & ~15];     /* 2^9 entries: the histogram (1272 B) sets the size */
    } u;
};
struct GroupLead { CodeScratch cs; uint32_t hist[QZ_HIST_WORDS]; };
struct GroupShared {
    uint32_t ticket, bfinal, btype, hb, pend, lead_buf;
    uint32_t nbytes[QZ_GROUP], ntok[QZ_GROUP], bits[QZ_GROUP], extra[QZ_GROUP];
};
template <int NT>
__device__ __forceinline__ void group_bar(uint32_t id)
{
#ifdef QZ_WARP_EMU
    emu::named_barrier(id, NT);
#else
    __syncwarp(); asm volatile("bar.sync %0, %1;" :: "r"(id), "n"(NT) : "memory");
#endif
}

/* take a free piece buffer: one bit per buffer in *busy (set = free); a warp that finds none sleeps with exponential
 * back-off instead of spinning on the issue slots the working warps need */
__device__ __forceinline__ uint32_t take_buffer(uint32_t *busy, uint32_t lane)
{
    uint32_t b = 0;
    if (lane == 0) {
        uint32_t ns = 128;
        for (;;) {
            const uint32_t m = *reinterpret_cast<volatile uint32_t *>(busy);
            if (m) {
                b = __ffs(m) - 1;
                if (atomicAnd(busy, ~(1u << b)) & (1u << b)) break;
                continue;
            }
            __nanosleep(ns);
            if (ns < 4096) ns <<= 1;
        }
        __threadfence_block();
    }
    return __shfl_sync(FULL, b, 0);
}
__device__ __forceinline__ void give_buffer(uint32_t *busy, uint32_t b, uint32_t lane)
{
    __syncwarp();
    if (lane == 0) { __threadfence_block(); atomicOr(busy, 1u << b); }
}

/* histogram and extra-bit total of one piece from its tokens in symbol form (mixed groups: the warp's own histogram
 * may cover two pieces) */
__device__ __forceinline__ uint32_t hist_from_tokens(uint32_t *hist, const uint32_t *toks, uint32_t ntok, uint32_t lane, uint64_t pkeep)
{
    for (uint32_t i = lane; i < QZ_HIST_WORDS; i += 32) hist[i] = 0;
    __syncwarp();
    uint32_t extra = 0;
    for (uint32_t i = lane; i < ntok; i += 32) {
        const uint32_t t = tok_ld(toks + i, pkeep);
        if (t & 0x80000000u) {
            const uint32_t ls = (t >> 26) & 31, ds = (t >> 16) & 31;
            atomicAdd(&hist[257 + ls], 1u);
            atomicAdd(&hist[QZ_DOFF + ds], 1u);
            extra += ((ls < 8 || ls == 28) ? 0u : (ls - 4) >> 2) + (ds < 4 ? 0u : (ds >> 1) - 1);
        } else atomicAdd(&hist[t & 511u], 1u);
    }
    __syncwarp();
    return warp_sum(extra);
}

/* GW warps per group, each with PPW = 8 / GW pieces of the block.  Only GW = 8 (one piece per warp) is instantiated: GW = 4
 * (two pieces per warp, so that three warps instead of seven wait for the leader) did cut that wait from 16 % to 11 % of warp
 * time on B200 but ran 30 % slower overall -- two pieces' tokens per warp no longer fit the L2 -- and is not offered. */
template <int PIECE_LOG2, int HB, int GW>
__global__ void __launch_bounds__(QZ_GROUPS_MAX_WARPS * 32) qzb_deflate_groups_kernel(QzbCompressJob job, int nbuf)
{
    constexpr int PIECE = 1 << PIECE_LOG2;
    constexpr int PPW = QZ_GROUP / GW;
    static_assert(sizeof(GroupWarpPriv<HB>) % 16 == 0 && sizeof(GroupWarpPriv<HB>) >= QZ_HIST_WORDS * 4, "a warp's slice holds its hash table, then its histogram");
    static_assert(sizeof(GroupLead) <= sizeof(PieceBuf<PIECE_LOG2>) && sizeof(PieceBuf<PIECE_LOG2>) % 16 == 0, "the code scratch borrows a piece buffer");
    QZ_DYN_SMEM(smem_raw);
    __shared__ uint32_t s_crc_tab[256];
    __shared__ uint32_t s_xstrip[5];
    __shared__ uint16_t s_lentab[256];
    __shared__ uint32_t s_busy[1];
    __shared__ GroupShared s_grp[QZ_GROUPS_MAX_WARPS / GW];
    constexpr uint32_t STRIP = PIECE / 32 + 4;

    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    PieceBuf<PIECE_LOG2> *bufs = reinterpret_cast<PieceBuf<PIECE_LOG2> *>(smem_raw);
    GroupWarpPriv<HB> *wsv = reinterpret_cast<GroupWarpPriv<HB> *>(smem_raw + (size_t)nbuf * sizeof(PieceBuf<PIECE_LOG2>));
    GroupWarpPriv<HB> &ws = wsv[warp];
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) { s_crc_tab[i] = qz_crc_table_entry(i); s_lentab[i] = len_table_entry(i); }
    if (threadIdx.x < 5) s_xstrip[threadIdx.x] = qz_crc_xpow8((uint64_t)STRIP << threadIdx.x);
    if (threadIdx.x == 0) s_busy[0] = nbuf >= 32 ? FULL : (1u << nbuf) - 1;
    __syncthreads();

    const uint32_t grp = warp / GW, wg = warp % GW, bar = 1 + grp;
    GroupShared &G = s_grp[grp];
    const uint32_t gwarp = blockIdx.x * nwarps + warp;
    uint32_t *toks0 = job.tok_scratch + (size_t)gwarp * PPW * QZB_TOK_STRIDE(PIECE);
    const uint32_t gpc = job.pieces_per_chunk / QZ_GROUP;     /* groups per chunk */
    const uint64_t pkeep = l2_policy_keep();
    uint32_t held = 0xffffffffu;                              /* leader: the piece buffer that serves as the block's code scratch */
#ifdef QZ_PHASE_CLOCKS
    long long tlast = clock64();
#endif
    for (;;) {
        if (wg == 0 && lane == 0) { G.ticket = atomicAdd(job.ticket, 1u); G.bfinal = 0; }
        group_bar<GW * 32>(bar);
        /* every warp of the group is past the previous block's emission: its code tables can go */
        if (wg == 0 && held != 0xffffffffu) { give_buffer(&s_busy[0], held, lane); held = 0xffffffffu; }
        const uint32_t gi = G.ticket;
        if (gi >= job.ngroups) break;
        const uint32_t chunk = gi / gpc, blk = gi - chunk * gpc;
        const uint32_t g0 = chunk * job.pieces_per_chunk + blk * QZ_GROUP;
        /* this warp's pieces: n = 0 for pieces behind the end of a ragged last chunk */
        PieceState ps[PPW];
        bool last_in_group[PPW];
#pragma unroll
        for (int j = 0; j < PPW; j++) {
            const uint32_t pi = wg * PPW + j;
            const uint64_t chunk_off = (uint64_t)chunk * job.chunk_sz;
            const uint64_t rem = job.src_len > chunk_off ? job.src_len - chunk_off : 0;
            const uint32_t chunk_len = rem < job.chunk_sz ? (uint32_t)rem : job.chunk_sz;
            const uint32_t p_off = (blk * QZ_GROUP + pi) << PIECE_LOG2;
            ps[j].g = g0 + pi; ps[j].n = chunk_len > p_off ? min((uint32_t)PIECE, chunk_len - p_off) : 0u; ps[j].ntok = 0; ps[j].extra_total = 0; ps[j].bfinal = false;
            ps[j].src = job.src + chunk_off + p_off;
            last_in_group[j] = ps[j].n != 0 && (pi == QZ_GROUP - 1 || p_off + ps[j].n == chunk_len);
        }
        /* phases 1-2 for every piece first (the hash table and the histogram share the warp's slice), then the token passes */
#pragma unroll
        for (int j = 0; j < PPW; j++) {
            if (ps[j].n) {
                const uint32_t b = take_buffer(&s_busy[0], lane);
                QZ_MARK(0);
                phase12<PIECE_LOG2, HB>(job, bufs[b].bytes, ws.u.table, toks0 + j * QZB_TOK_STRIDE(PIECE), s_crc_tab, s_xstrip, ps[j].g, lane, ps[j] QZ_TPASS);
                give_buffer(&s_busy[0], b, lane);
            }
        }
        uint32_t extra = 0;
#pragma unroll
        for (int j = 0; j < PPW; j++) {
            uint32_t *toks = toks0 + j * QZB_TOK_STRIDE(PIECE);
            extra += token_pass(ws.u.hist, toks, ps[j].ntok, s_lentab, lane, pkeep, j == 0);
            /* the last piece of the block carries the end-of-block token */
            if (lane == 0 && last_in_group[j]) tok_st(toks + ps[j].ntok, 256u, pkeep);
        }
        extra = warp_sum(extra);
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < PPW; j++) {
                G.nbytes[wg * PPW + j] = ps[j].n; G.ntok[wg * PPW + j] = ps[j].ntok + (last_in_group[j] ? 1u : 0u);
                if (ps[j].bfinal) G.bfinal = 1;
            }
            G.extra[wg] = extra;
        }
        QZ_MARK(3);
        group_bar<GW * 32>(bar);
        QZ_MARK(9);                 /* waiting for the group's slowest piece */
        /* A group that mixes incompressible pieces (close to one token per byte) with compressible ones is better off
         * with a block per piece: one code table cannot serve both, and only whole blocks can fall back to stored. */
        {
            uint32_t hi = 0, lo = 0xffffffffu;
#pragma unroll
            for (int i = 0; i < QZ_GROUP; i++) {
                const uint32_t nb = G.nbytes[i];
                if (nb) { const uint32_t r = (G.ntok[i] << 10) / nb; hi = max(hi, r); lo = min(lo, r); }
            }
            if (hi > 920u && lo < 768u) {
                /* rare: every piece on its own, with a piece buffer as its code scratch */
#pragma unroll
                for (int j = 0; j < PPW; j++) {
                    if (!ps[j].n) continue;
                    uint32_t *toks = toks0 + j * QZB_TOK_STRIDE(PIECE);
                    if (lane == 0) tok_st(toks + ps[j].ntok, 256u, pkeep);
                    const uint32_t b = take_buffer(&s_busy[0], lane);
                    GroupLead &own = *reinterpret_cast<GroupLead *>(bufs[b].bytes);
                    const uint32_t ex = hist_from_tokens(own.hist, toks, ps[j].ntok, lane, pkeep);
                    finish_piece(job, own.cs, own.hist, toks, lane, ps[j], ex, pkeep QZ_TPASS);
                    give_buffer(&s_busy[0], b, lane);
                }
                continue;
            }
        }
        /* leader: one histogram, one set of codes, one block header for the group */
        if (wg == 0) {
            held = take_buffer(&s_busy[0], lane);
            GroupLead &L = *reinterpret_cast<GroupLead *>(bufs[held].bytes);
            uint32_t extra_total = 0, nbytes = 0, npc = 0;
            for (int i = 0; i < GW; i++) extra_total += G.extra[i];
            for (int i = 0; i < QZ_GROUP; i++) { nbytes += G.nbytes[i]; npc += G.nbytes[i] ? 1u : 0u; }
            for (uint32_t s = lane; s < QZ_HIST_WORDS; s += 32) {
                uint32_t f = 0;
#pragma unroll
                for (int i = 0; i < GW; i++) f += wsv[grp * GW + i].u.hist[s];
                L.hist[s] = f;
            }
            __syncwarp();
            const int btype = choose_block(L.cs, L.hist, extra_total, (5 * npc + nbytes) * 8, job.static_huffman, lane QZ_TPASS);
            uint32_t hb = 0, pend = 0;
            if (btype) open_block(L.cs, L.hist, btype, G.bfinal != 0, reinterpret_cast<uint32_t *>(job.slots + (size_t)g0 * job.slot_stride), lane, &hb, &pend QZ_TPASS);
            if (lane == 0) { G.btype = (uint32_t)btype; G.hb = hb; G.pend = pend; G.lead_buf = held; }
            QZ_MARK(11);            /* leader: block header written, tables final */
        }
        group_bar<GW * 32>(bar);
        QZ_MARK(10);                /* waiting for the leader */
        const uint32_t btype = G.btype;
        const bool gfinal = G.bfinal != 0;
        if (btype == 0) {
            /* incompressible group: every piece is its own stored block in its own slot, as in the per-piece kernel */
#pragma unroll
            for (int j = 0; j < PPW; j++) {
                if (!ps[j].n) continue;
                const uint32_t out_bytes = stored_piece(job.slots + (size_t)ps[j].g * job.slot_stride, ps[j].src, ps[j].n, ps[j].bfinal, lane);
                if (lane == 0) job.piece_len[ps[j].g] = out_bytes;
            }
        } else {
            const uint32_t *tab = reinterpret_cast<const GroupLead *>(bufs[G.lead_buf].bytes)->hist;
            uint32_t *slotw = reinterpret_cast<uint32_t *>(job.slots + (size_t)g0 * job.slot_stride);
            uint32_t beg[PPW], end[PPW], mybits[PPW], incl[PPW];
#pragma unroll
            for (int j = 0; j < PPW; j++) {
                const uint32_t NT = G.ntok[wg * PPW + j];
                const uint32_t R = (((NT + 31) >> 5) + 3) & ~3u;
                beg[j] = min(lane * R, NT); end[j] = min(beg[j] + R, NT);
                mybits[j] = count_run_bits(tab, toks0 + j * QZB_TOK_STRIDE(PIECE), beg[j], end[j], pkeep);
                incl[j] = mybits[j];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(FULL, incl[j], o); if (lane >= (uint32_t)o) incl[j] += y; }
                if (lane == 31) G.bits[wg * PPW + j] = incl[j];
            }
            QZ_MARK(12);            /* count pass */
            group_bar<GW * 32>(bar);
            QZ_MARK(13);
            uint32_t before[PPW], total = G.hb;
#pragma unroll
            for (int j = 0; j < PPW; j++) before[j] = G.hb;
#pragma unroll
            for (int i = 0; i < QZ_GROUP; i++) {
                const uint32_t bi = G.bits[i];
#pragma unroll
                for (int j = 0; j < PPW; j++) if (i < (int)(wg * PPW + j)) before[j] += bi;
                total += bi;
            }
            const uint32_t end_bit = total;
            const uint32_t nz = 3 + ((0u - (end_bit + 3)) & 7);
            const uint32_t end_bit2 = gfinal ? end_bit : end_bit + nz + 32;
#pragma unroll
            for (int j = 0; j < PPW; j++) slotw[(before[j] + incl[j] - mybits[j]) >> 5] = 0;
            if (wg == GW - 1 && lane == 31) slotw[end_bit2 >> 5] = 0;
            group_bar<GW * 32>(bar);
            QZ_MARK(14);
#pragma unroll
            for (int j = 0; j < PPW; j++) {
                const uint32_t NT = G.ntok[wg * PPW + j];
                /* the lane that codes the end-of-block token appends the trailer; it is the last token of the group */
                const bool owns_eob = last_in_group[j] && beg[j] < NT && end[j] == NT;
                emit_run(tab, toks0 + j * QZB_TOK_STRIDE(PIECE), beg[j], end[j], before[j] + incl[j] - mybits[j], G.pend, wg == 0 && j == 0 && lane == 0,
                         !gfinal && owns_eob, nz, slotw, pkeep);
                if (lane == 0 && ps[j].n) job.piece_len[ps[j].g] = (wg == 0 && j == 0) ? (end_bit2 + 7) >> 3 : 0u;        /* pieces behind a ragged end have no entry */
            }
        }
        __syncwarp();
        QZ_MARK(8);
    }
}

#ifdef QZ_SPLIT_KERNEL
#include "qz_deflate_split.cuh"
#endif

/* ------------------------------------------------------------------------------------------ */
/* Framing: sizes -> exclusive scan -> headers, payload gather, footers.
 * Replaces doCompressOut's per-chunk header gen / payload memcpy / crc32_combine / footer gen
 * (reference src/qatzip.c:1699-1716, src/qatzip_gzip.c:98-143,228-237, src/qatzip_lz4.c:104-143). */

__device__ __forceinline__ uint32_t qzb_hdr_sz(int fmt) { return fmt == QZB_FMT_GZIP_EXT ? 24u : fmt == QZB_FMT_GZIP ? 10u : fmt == QZB_FMT_4B ? 4u : fmt == QZB_FMT_LZ4 ? 15u : fmt == QZB_FMT_ZLIB ? 2u : 0u; }
__device__ __forceinline__ uint32_t qzb_ftr_sz(int fmt) { return (fmt == QZB_FMT_GZIP_EXT || fmt == QZB_FMT_GZIP || fmt == QZB_FMT_LZ4) ? 8u : fmt == QZB_FMT_ZLIB ? 4u : 0u; }

__global__ void qzb_chunk_sizes_kernel(QzbCompressJob job)
{
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= job.nchunks) return;
    uint32_t g0 = c * job.pieces_per_chunk, g1 = min(g0 + job.pieces_per_chunk, job.npieces), payload = 0;
    for (uint32_t g = g0; g < g1; g++) payload += job.piece_len[g];
    job.chunk_total[c] = qzb_hdr_sz(job.fmt) + payload + qzb_ftr_sz(job.fmt);
}

/* single-CTA exclusive scan of chunk_total into chunk_off[0..nchunks] */
__global__ void __launch_bounds__(1024) qzb_scan_kernel(const uint32_t *in, uint64_t *out, uint32_t n)
{
    __shared__ uint64_t s_warp[32];
    __shared__ uint64_t s_carry;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        uint32_t i = base + threadIdx.x;
        uint64_t v = i < n ? in[i] : 0, incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint64_t y = __shfl_up_sync(FULL, incl, o); if (lane >= (uint32_t)o) incl += y; }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint64_t w = s_warp[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { uint64_t y = __shfl_up_sync(FULL, wi, o); if (lane >= (uint32_t)o) wi += y; }
            s_warp[lane] = wi - w;
        }
        __syncthreads();
        uint64_t excl = s_carry + s_warp[warp] + incl - v;
        if (i < n) out[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = s_carry;
}

#include "qz_xxh32.h"
__device__ __forceinline__ void st32le(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }

/* One CTA (8 warps) per chunk.  Warp 0 turns the pieces' lengths into offsets and, lane per piece, their
 * checksums into the chunk's (every piece's CRC-32 times x^(8 * bytes after it), XOR-ed: the pieces' terms
 * are independent); then every warp copies whole pieces, 4 bytes per lane with the loads of eight steps in
 * flight, realigning through a funnel shift when the destination is not word-aligned. */
#define QZ_FRAME_WARPS 8
__global__ void __launch_bounds__(QZ_FRAME_WARPS * 32) qzb_frame_kernel(QzbCompressJob job)
{
    const uint32_t c = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t off = job.chunk_off[c];
    const uint32_t total = job.chunk_total[c];
    if (off + total > job.dst_cap) return;             /* host reports QZ_BUF_ERROR for this and later chunks */
    const uint32_t g0 = c * job.pieces_per_chunk, g1 = min(g0 + job.pieces_per_chunk, job.npieces), np = g1 - g0;   /* np <= 64 */
    const uint64_t chunk_in = (uint64_t)c * job.chunk_sz;
    const uint64_t rem = job.src_len > chunk_in ? job.src_len - chunk_in : 0;
    const uint32_t chunk_len = rem < job.chunk_sz ? (uint32_t)rem : job.chunk_sz;
    const uint32_t hs = qzb_hdr_sz(job.fmt), fs = qzb_ftr_sz(job.fmt), payload = total - hs - fs;
    const uint32_t PIECE = 1u << job.piece_log2;
    uint8_t *__restrict__ d = job.dst + off;
    __shared__ uint32_t s_poff[65];                      /* exclusive prefix of the pieces' lengths */
    if (warp == 0) {
        /* lengths -> offsets: two pieces per lane */
        const uint32_t l0 = lane < np ? job.piece_len[g0 + lane] : 0u, l1 = lane + 32 < np ? job.piece_len[g0 + 32 + lane] : 0u;
        uint32_t i0 = l0, i1 = l1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y0 = __shfl_up_sync(FULL, i0, o), y1 = __shfl_up_sync(FULL, i1, o); if (lane >= (uint32_t)o) { i0 += y0; i1 += y1; } }
        const uint32_t t0 = __shfl_sync(FULL, i0, 31);
        s_poff[lane] = i0 - l0; s_poff[32 + lane] = t0 + i1 - l1;
        if (lane == 31) s_poff[64] = t0 + i1;
        /* checksum of the chunk */
        uint32_t ck;
        if (job.fmt == QZB_FMT_LZ4) ck = job.chunk_cksum[c];      /* XXH32 written by the xxh kernel */
        else if (job.fmt == QZB_FMT_ZLIB) {
            /* per-piece Adler sums -> chunk Adler-32 (reference footer: src/qatzip_gzip.c:273-281) */
            ck = 0;
            if (lane == 0) {
                uint32_t s1 = 0, s2 = 0;
                for (uint32_t g = g0; g < g1; g++) {
                    const uint32_t pn = min(PIECE, chunk_len - (g - g0) * PIECE), w = job.piece_crc[g];
                    qz_adler_join(&s1, &s2, w & 0xffffu, w >> 16, pn);
                }
                ck = qz_adler_finish(s1, s2, chunk_len);
            }
            ck = __shfl_sync(FULL, ck, 0);
        } else {
            ck = 0;
            for (uint32_t k = lane; k < np; k += 32) {
                const uint32_t after = chunk_len - min(chunk_len, (k + 1) * PIECE);      /* input bytes behind piece k */
                ck ^= qz_gf2_mul(job.piece_crc[g0 + k], qz_crc_xpow8(after));
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) ck ^= __shfl_xor_sync(FULL, ck, o);
        }
        if (lane == 0) {
            if (job.fmt != QZB_FMT_LZ4) job.chunk_cksum[c] = ck;
            switch (job.fmt) {
            case QZB_FMT_GZIP_EXT:
                d[10] = 12; d[11] = 0; d[12] = 'Q'; d[13] = 'Z'; d[14] = 8; d[15] = 0;
                st32le(d + 16, chunk_len); st32le(d + 20, payload);
                /* fall through */
            case QZB_FMT_GZIP:
                d[0] = 0x1f; d[1] = 0x8b; d[2] = 8; d[3] = (job.fmt == QZB_FMT_GZIP_EXT) ? 4 : 0;
                d[4] = d[5] = d[6] = d[7] = 0; d[8] = 0; d[9] = 0xff;
                break;
            case QZB_FMT_4B: st32le(d, payload); break;
            case QZB_FMT_ZLIB: d[0] = 0x78; d[1] = 0x9C; break;           /* reference src/qatzip_gzip.c:263-271 */
            case QZB_FMT_LZ4:
                st32le(d, 0x184D2204u); d[4] = 0x4C; d[5] = 0x40; st32le(d + 6, chunk_len); st32le(d + 10, 0);
                d[14] = (uint8_t)(qz_xxh32(d + 4, 10, 0) >> 8);
                break;
            default: break;
            }
            if (fs) {
                uint8_t *f = d + hs + payload;
                if (job.fmt == QZB_FMT_LZ4) { st32le(f, 0); st32le(f + 4, ck); }
                else if (job.fmt == QZB_FMT_ZLIB) { f[0] = (uint8_t)(ck >> 24); f[1] = (uint8_t)(ck >> 16); f[2] = (uint8_t)(ck >> 8); f[3] = (uint8_t)ck; }
                else { st32le(f, ck); st32le(f + 4, chunk_len); }
            }
        }
    }
    __syncthreads();
    /* payload gather: every piece's bytes (a group's block sits whole in its first piece's slot, the other seven are
     * empty) are cut into runs of 4 KiB that are dealt to the warps in turn */
    constexpr uint32_t SUB = 4096;
    uint32_t item = 0;
    for (uint32_t k = 0; k < np; k++) {
        const uint32_t plen = s_poff[k + 1] - s_poff[k];
        for (uint32_t o = 0; o < plen; o += SUB, item++) {
            if (item % QZ_FRAME_WARPS != warp) continue;
            const uint32_t len = min(SUB, plen - o);
            const uint8_t *__restrict__ s = job.slots + (size_t)(g0 + k) * job.slot_stride + o;     /* slot and o are 16-byte aligned */
            uint8_t *__restrict__ dd = d + hs + s_poff[k] + o;
            const uint32_t mis = (uint32_t)((4 - (reinterpret_cast<uintptr_t>(dd) & 3)) & 3);
            const uint32_t head = min(mis, len);
            if (lane < head) dd[lane] = s[lane];
            const uint32_t nw = (len - head) >> 2;
            uint32_t *__restrict__ dw = reinterpret_cast<uint32_t *>(dd + head);
            const uint32_t *__restrict__ sw = reinterpret_cast<const uint32_t *>(s);
            const uint32_t sh = head * 8;
            for (uint32_t i0 = lane; i0 < nw; i0 += 32 * 8) {
                uint32_t a[8], b2[8];
#pragma unroll
                for (int u = 0; u < 8; u++) { const uint32_t i = i0 + 32 * u; a[u] = i < nw ? sw[i] : 0u; b2[u] = (sh && i < nw) ? sw[i + 1] : 0u; }
#pragma unroll
                for (int u = 0; u < 8; u++) { const uint32_t i = i0 + 32 * u; if (i < nw) dw[i] = sh ? __funnelshift_r(a[u], b2[u], sh) : a[u]; }
            }
            for (uint32_t i = head + (nw << 2) + lane; i < len; i += 32) dd[i] = s[i];
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
#ifndef QZ_WARP_EMU
/* shared memory for `warps` warps sharing `nbuf` piece buffers */
extern "C" size_t qzb_deflate_smem_bytes(int piece_log2, int hb, int warps, int nbuf)
{
    size_t priv = hb == 11 ? sizeof(WarpPriv<11>) : hb == 12 ? sizeof(WarpPriv<12>) : sizeof(WarpPriv<13>);
    size_t buf = piece_log2 == 13 ? sizeof(PieceBuf<13>) : sizeof(PieceBuf<14>);
    return priv * (size_t)warps + buf * (size_t)nbuf;
}

template <int P, int H>
static cudaError_t launch_deflate(const QzbCompressJob &job, int grid, int warps, int nbuf, cudaStream_t st)
{
    size_t smem = sizeof(WarpPriv<H>) * (size_t)warps + sizeof(PieceBuf<P>) * (size_t)nbuf;
    cudaError_t e = cudaFuncSetAttribute(qzb_deflate_pieces_kernel<P, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    qzb_deflate_pieces_kernel<P, H><<<grid, warps * 32, smem, st>>>(job, nbuf);
    return cudaGetLastError();
}

extern "C" cudaError_t qzb_launch_deflate(const QzbCompressJob *job, int hb, int grid, int warps, int nbuf, cudaStream_t st)
{
    if (nbuf < 1 || nbuf > 32 || warps < 1 || warps > QZ_PIECES_MAX_WARPS) return cudaErrorInvalidValue;
    if (job->piece_log2 == 13 && hb == 11) return launch_deflate<13, 11>(*job, grid, warps, nbuf, st);
    if (job->piece_log2 == 13 && hb == 12) return launch_deflate<13, 12>(*job, grid, warps, nbuf, st);
    if (job->piece_log2 == 14 && hb == 12) return launch_deflate<14, 12>(*job, grid, warps, nbuf, st);
    if (job->piece_log2 == 14 && hb == 13) return launch_deflate<14, 13>(*job, grid, warps, nbuf, st);
    return cudaErrorInvalidValue;
}

extern "C" int qzb_deflate_max_warps(int group) { return group ? QZ_GROUPS_MAX_WARPS : QZ_PIECES_MAX_WARPS; }

/* shared memory of the group kernel for `warps` warps sharing `nbuf` piece buffers (the code scratch borrows from the pool) */
extern "C" size_t qzb_deflate_groups_smem_bytes(int hb, int warps, int nbuf)
{
    const size_t priv = hb == 9 ? sizeof(GroupWarpPriv<9>) : (size_t)2 << hb;
    return priv * (size_t)warps + sizeof(PieceBuf<13>) * (size_t)nbuf;
}

template <int P, int H, int GW>
static cudaError_t launch_deflate_groups(const QzbCompressJob &job, int grid, int warps, int nbuf, cudaStream_t st)
{
    size_t smem = qzb_deflate_groups_smem_bytes(H, warps, nbuf);
    cudaError_t e = cudaFuncSetAttribute(qzb_deflate_groups_kernel<P, H, GW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    qzb_deflate_groups_kernel<P, H, GW><<<grid, warps * 32, smem, st>>>(job, nbuf);
    return cudaGetLastError();
}

/* group kernel (one deflate block per QZ_GROUP pieces): warps a multiple of QZ_GROUP, job->ngroups set */
extern "C" cudaError_t qzb_launch_deflate_groups(const QzbCompressJob *job, int hb, int grid, int warps, int nbuf, cudaStream_t st)
{
    if (nbuf < 1 || nbuf > 32 || warps < QZ_GROUP || warps > QZ_GROUPS_MAX_WARPS || warps % QZ_GROUP || job->pieces_per_chunk % QZ_GROUP || !job->ngroups || job->piece_log2 != 13)
        return cudaErrorInvalidValue;
    if (hb == 9) return launch_deflate_groups<13, 9, 8>(*job, grid, warps, nbuf, st);
    if (hb == 10) return launch_deflate_groups<13, 10, 8>(*job, grid, warps, nbuf, st);
    if (hb == 11) return launch_deflate_groups<13, 11, 8>(*job, grid, warps, nbuf, st);
    if (hb == 12) return launch_deflate_groups<13, 12, 8>(*job, grid, warps, nbuf, st);
    return cudaErrorInvalidValue;
}

/* experimental matcher / coder kernel (qz_deflate_split.cuh): present only in builds with -DQZ_SPLIT_KERNEL */
#ifdef QZ_SPLIT_KERNEL
extern "C" int qzb_deflate_split_compiled(void) { return 1; }
extern "C" size_t qzb_deflate_split_smem_bytes(int hb, int nmatch, int nteams) { return qzs_smem_bytes(hb, nmatch, nteams); }
extern "C" size_t qzb_deflate_split_tok_words(int grid) { return (size_t)grid * QZS_SLOTS * QZ_GROUP * QZB_TOK_STRIDE(1 << 13); }
template <int H>
static cudaError_t launch_deflate_split(const QzbCompressJob &job, int grid, int nmatch, int nteams, cudaStream_t st)
{
    const size_t smem = qzs_smem_bytes(H, nmatch, nteams);
    cudaError_t e = cudaFuncSetAttribute(qzb_deflate_split_kernel<13, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    qzb_deflate_split_kernel<13, H><<<grid, (nmatch + nteams * QZS_TEAM) * 32, smem, st>>>(job, nmatch, nteams);
    return cudaGetLastError();
}
extern "C" cudaError_t qzb_launch_deflate_split(const QzbCompressJob *job, int hb, int grid, int nmatch, int nteams, cudaStream_t st)
{
    if (nmatch < 1 || nteams < 1 || nteams > 8 || nmatch + nteams * QZS_TEAM > 32 || job->pieces_per_chunk % QZ_GROUP || !job->ngroups || job->piece_log2 != 13) return cudaErrorInvalidValue;
    if (hb == 10) return launch_deflate_split<10>(*job, grid, nmatch, nteams, st);
    if (hb == 11) return launch_deflate_split<11>(*job, grid, nmatch, nteams, st);
    return cudaErrorInvalidValue;
}
#else
extern "C" int qzb_deflate_split_compiled(void) { return 0; }
extern "C" size_t qzb_deflate_split_smem_bytes(int, int, int) { return 0; }
extern "C" size_t qzb_deflate_split_tok_words(int) { return 0; }
extern "C" cudaError_t qzb_launch_deflate_split(const QzbCompressJob *, int, int, int, int, cudaStream_t) { return cudaErrorNotSupported; }
#endif

extern "C" cudaError_t qzb_launch_frame(const QzbCompressJob *job, cudaStream_t st)
{
    qzb_chunk_sizes_kernel<<<(job->nchunks + 255) / 256, 256, 0, st>>>(*job);
    qzb_scan_kernel<<<1, 1024, 0, st>>>(job->chunk_total, job->chunk_off, job->nchunks);
    qzb_frame_kernel<<<job->nchunks, QZ_FRAME_WARPS * 32, 0, st>>>(*job);
    return cudaGetLastError();
}
#endif
