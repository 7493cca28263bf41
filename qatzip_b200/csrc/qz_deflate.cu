/* qz_deflate.cu -- sm_100a DEFLATE compressor.
 *
 * These kernels replace the QAT compress request submitted at reference src/qatzip.c:1542 (cpaDcCompressData2, stateless
 * deflate, CPA_DC_FLUSH_FINAL / _FULL) with the session set-up of reference src/qatzip_utils.c:264-341 (32 KiB history,
 * dynamic or static Huffman, stored fallback, CRC-32 of the input returned in res.checksum), and the stitching
 * doCompressOut does afterwards (framing kernels at the end of the file).
 *
 * The device's units: a PIECE is 8 KiB of a chunk, matched by one warp; a WINDOW is 64 KiB of a chunk (eight pieces) that
 * sits whole in shared memory while its pieces are matched, so every position sees the whole window in front of it
 * (qz_match.cuh), and that becomes one deflate block.  HBM sees each input byte once and each output byte once; tokens
 * live as 16-bit slots in an L2-resident scratch in between.
 *
 * Stages of a piece, all warp-synchronous:
 *   1 load    global -> shared, 16 B per lane, then CRC-32 / Adler-32 over right-aligned per-lane strips   (load_and_checksum)
 *   2 match   qz_match.cuh: 32 positions per step, single-probe 4-byte hash + in-tile candidates, 12-byte verify in the
 *             lane, greedy parse by pointer doubling, slots to the scratch
 *   3 code    histogram of the slots; sort by frequency in registers, in-place Huffman lengths (merge on one lane,
 *             depths by pointer doubling, leaf depths across the warp), canonical codes, dynamic header plan; cheapest of
 *             stored / fixed / dynamic                                       (slot_hist, choose_block, open_block)
 *   4 emit    every lane packs a contiguous run of slots at a bit offset from a scan of the runs' lengths; run-boundary
 *             words are zeroed first and joined by atomic OR                          (count_run_bits, emit_run)
 *
 * Two kernels string them together:
 *   qzb_deflate_window_kernel  (hw_buff_sz >= 64 KiB) eight warps = one window = ONE deflate block: stages 1-2 and the
 *                              histogram per warp, the code construction once per window by the group's leader,
 *                              emission at bit offsets inside the window's output
 *   qzb_deflate_pieces_kernel  smaller chunks: one block per piece with a private window, no cross-warp step
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include "qz_kernels.cuh"
#include "qz_warp.cuh"
#include "qz_match.cuh"
#include "qz_huffman.h"
#include "qz_crc32.h"
#include "qz_adler32.h"

#define FULL 0xffffffffu
#define QZ_STAGE_WORDS 64
/* most warps a CTA may have.  Per-piece kernel: shared memory admits 20-24, and the bound lets the compiler use up to
 * 80 registers per thread instead of the 64 a 1024-thread bound would impose.  Window kernel: four groups of eight warps
 * share two units. */
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

/* Everything the code construction and the emission keep in shared memory for one block: sort keys -> frequencies ->
 * header-plan counters -> bit staging window -> (once the block header is out) the 256 per-length code entries; symbol
 * ids, code lengths, the dynamic header; and the histograms that become the code tables.  4000 B. */
struct CodeScratch {
    uint32_t keys[288];                 /* 286 sort keys at most; see above for its later lives */
    uint16_t ids[QZ_NUM_LL + 2];
    uint8_t ll_len[288];
    uint8_t d_len[32];
    QzDynHeaderCore hdr;
};
#define QZ_HIST_WORDS (QZ_NUM_LL + 2 + QZ_NUM_D + 2)   /* [0,286) lit/len, [288,318) dist; later the code tables */
#define QZ_DOFF 288
struct BlockCoder { CodeScratch cs; uint32_t hist[QZ_HIST_WORDS]; };
/* code entry of length slot l (len - 3), relative to the code tables in BlockCoder::hist: it lives in cs.keys */
#define QZ_LENVAL_OFF (-(int)(sizeof(CodeScratch) / 4))

/* Per-piece kernel: a CTA owns NB piece buffers and NW > NB warps.  A warp draws a ticket, takes any free buffer, loads and
 * matches its piece (private window: the warp's slice is the hash table), hands the buffer back and codes the piece from its
 * slots in the L2 scratch, the slice now being the block coder. */
template <int HB>
struct WarpPriv {
    union {
        uint16_t table[1 << HB];
        BlockCoder b;
    } u;
};
template <int PIECE_LOG2>
struct PieceBuf {
    uint8_t bytes[QZM_FRONT_PAD + (1 << PIECE_LOG2) + QZM_TAIL_PAD];     /* data at bytes + QZM_FRONT_PAD */
};

/* what the coding stage needs to know about a matched piece */
struct PieceState {
    uint32_t g, n, nslots;
    bool bfinal;
    const uint8_t *src;
};

/* ---- stage 1: global -> shared, 16 B per lane, and the checksum of the piece (CRC-32, or packed Adler sums for zlib
 * streams) over right-aligned per-lane strips joined up a tree.  `pad`: zero the bytes behind the data. */
template <int PIECE>
__device__ __forceinline__ uint32_t load_and_checksum(uint8_t *piece, const uint8_t *src, uint32_t n, bool pad, int fmt,
                                                      const uint32_t *s_crc_tab, const uint32_t *s_xstrip, uint32_t lane)
{
    constexpr uint32_t STRIP = PIECE / 32 + 4;   /* bytes per lane; /4 is odd -> conflict-free banks */
    uint4 *d4 = reinterpret_cast<uint4 *>(piece);
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
        const uint32_t nv = n >> 4;
        const uint64_t pstream = l2_policy_stream();
        for (uint32_t i = lane; i < nv; i += 32) d4[i] = stream_ld16(s4 + i, pstream);
        for (uint32_t i = (nv << 4) + lane; i < n; i += 32) piece[i] = src[i];
    } else {
        for (uint32_t i = lane; i < n; i += 32) piece[i] = src[i];
    }
    if (pad) { piece[n + lane] = 0; if (lane < QZM_TAIL_PAD - 32) piece[n + 32 + lane] = 0; }
    __syncwarp();
    /* right-aligned strips: lane i owns [n-(32-i)*STRIP, n-(31-i)*STRIP) clipped at 0 */
    int hi = (int)n - (int)((31 - lane) * STRIP), lo = hi - (int)STRIP;
    if (lo < 0) lo = 0;
    uint32_t c;
    if (fmt == QZB_FMT_ZLIB) {
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
    return c;           /* valid in lane 0 */
}

/* geometry of piece g of the job */
template <int PIECE_LOG2>
__device__ __forceinline__ void piece_geometry(const QzbCompressJob &job, uint32_t g, PieceState &ps)
{
    constexpr int PIECE = 1 << PIECE_LOG2;
    const uint32_t chunk = g / job.pieces_per_chunk, k = g - chunk * job.pieces_per_chunk;
    const uint64_t chunk_off = (uint64_t)chunk * job.chunk_sz;
    const uint64_t rem = job.src_len > chunk_off ? job.src_len - chunk_off : 0;
    const uint32_t chunk_len = rem < job.chunk_sz ? (uint32_t)rem : job.chunk_sz;
    const uint32_t p_off = k << PIECE_LOG2;
    const uint32_t n = chunk_len > p_off ? min((uint32_t)PIECE, chunk_len - p_off) : 0u;
    const bool last_piece = (p_off + n == chunk_len);
    ps.g = g; ps.n = n; ps.nslots = 0;
    ps.bfinal = last_piece && (job.fmt != QZB_FMT_RAW || (chunk == job.nchunks - 1 && job.last));
    ps.src = job.src + chunk_off + p_off;
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

/* ---- stage 3a: histograms of a run of slots (every slot is one code: literal / end of block, length, or distance) ----
 * Returns the lane's share of the extra bits. */
__device__ __forceinline__ uint32_t slot_hist(uint32_t *hist, const uint16_t *slots, uint32_t nslots, const uint16_t *s_lentab, uint32_t lane, uint64_t pkeep)
{
    uint32_t extra = 0;
    for (uint32_t i0 = lane * 8; i0 < nslots; i0 += 256) {
        const uint4 q = tok_ld4(reinterpret_cast<const uint32_t *>(slots + i0), pkeep);
        const uint32_t w[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (i0 + k < nslots) {
                /* one path for the three slot kinds (no divergence): symbol index and extra-bit count by selects */
                const uint32_t s = (w[k >> 1] >> ((k & 1) * 16)) & 0xffffu;
                const bool isD = (s & QZ_SLOT_DIST) != 0, isL = !isD && (s & QZ_SLOT_LEN) != 0;
                const uint32_t x = s & 0x7fffu;                 /* distance - 1 */
                const uint32_t lg = 31 - __clz((int)(x | 1u));
                const uint32_t de = x < 4 ? 0u : lg - 1, ds = x < 4 ? x : 2 * lg + ((x >> de) & 1);
                const uint32_t le = s_lentab[s & 0xff];
                atomicAdd(&hist[isD ? QZ_DOFF + ds : isL ? 257 + (le & 31) : s], 1u);
                extra += isD ? de : isL ? (le >> 5) & 7 : 0u;
            }
        }
    }
    return extra;
}

/* stage 3b: code construction from the histograms in hist -> code lengths in cs.ll_len / cs.d_len, the planned
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
 * were, the block header is written; *hb = bits written so far, *pend = the
 * partial word at that position (the first token run continues it) */
__device__ __forceinline__ void open_block(CodeScratch &cs, uint32_t *hist, int btype, bool bfinal, uint32_t *slotw, const uint16_t *s_lentab, uint32_t lane, uint32_t *hb, uint32_t *pend_out QZ_TARG)
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
    /* The code tables take their final form, one entry per slot kind, value | bit count << 24:
     *   literal / end of block  hist[s]                 code
     *   length slot l           hist[QZ_LENVAL_OFF + l] length code with its extra bits appended (in cs.keys: the staging window is done)
     *   distance symbol         hist[QZ_DOFF + ds]      code | code length << 16 | (code length + extra-bit count) << 24 */
    __syncwarp();           /* every lane has read the pending word out of the staging window */
    {
        uint32_t *lenval = hist + QZ_LENVAL_OFF;
        for (uint32_t l = lane; l < 256; l += 32) {
            const uint32_t le = s_lentab[l], c = hist[257 + (le & 31)], clen = (c >> 16) & 0xff;
            lenval[l] = (c & 0xffffu) | ((l - (le >> 8)) << clen) | ((clen + ((le >> 5) & 7)) << 24);
        }
        __syncwarp();
        for (uint32_t s2 = lane; s2 < 257; s2 += 32) { const uint32_t c = hist[s2]; hist[s2] = (c & 0xffffu) | (((c >> 16) & 0xff) << 24); }
        if (lane < QZ_NUM_D) { const uint32_t c = hist[QZ_DOFF + lane], clen = (c >> 16) & 0xff; hist[QZ_DOFF + lane] = (c & 0xffffu) | (clen << 16) | ((clen + (lane < 4 ? 0u : (lane >> 1) - 1)) << 24); }
    }
    __syncwarp();
}

/* ---- stage 4: emission.  A lane codes a contiguous run of slots ---- */
/* slot -> (bits, bit count) under the code tables `tab` */
__device__ __forceinline__ void slot_code(const uint32_t *tab, uint32_t s, uint32_t &val, uint32_t &nb)
{
    const bool isD = (s & QZ_SLOT_DIST) != 0;
    const uint32_t x = s & 0x7fffu;                 /* distance - 1 */
    const uint32_t lg = 31 - __clz((int)(x | 1u));
    const uint32_t de = x < 4 ? 0u : lg - 1, ds = x < 4 ? x : 2 * lg + ((x >> de) & 1), dv = x & ((1u << de) - 1);
    const int idx = isD ? QZ_DOFF + (int)ds : (s & QZ_SLOT_LEN) ? QZ_LENVAL_OFF + (int)(s & 0xff) : (int)s;
    const uint32_t e = tab[idx];
    nb = e >> 24;
    val = isD ? (e & 0xffffu) | (dv << ((e >> 16) & 0xff)) : (e & 0xffffffu);
}

/* pass 1: bits the slots [beg, end) take (beg a multiple of 8: the slots are fetched 16 bytes at a time) */
__device__ __forceinline__ uint32_t count_run_bits(const uint32_t *tab, const uint16_t *slots, uint32_t beg, uint32_t end, uint64_t pkeep)
{
    uint32_t mybits = 0;
    uint4 qn = beg < end ? tok_ld4(reinterpret_cast<const uint32_t *>(slots + beg), pkeep) : make_uint4(0, 0, 0, 0);   /* one group ahead: hides the L2 round trip */
    for (uint32_t j = beg; j < end; j += 8) {
        const uint32_t w[4] = { qn.x, qn.y, qn.z, qn.w };
        if (j + 8 < end) qn = tok_ld4(reinterpret_cast<const uint32_t *>(slots + j + 8), pkeep);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (j + k < end) {
                uint32_t val, nb;
                slot_code(tab, (w[k >> 1] >> ((k & 1) * 16)) & 0xffffu, val, nb);
                mybits += nb;
            }
        }
    }
    return mybits;
}

/* pass 2: pack the slots [beg, end) at bit `start` of slotw through a private 64-bit accumulator.  `acc0` is what already
 * sits in the first word below the start bit and `owns_first` says this lane writes that word whole (it continues the block
 * header); every other lane joins its first word by atomic OR when it starts inside one.  With `trailer` the lane appends the
 * byte-aligning empty stored block (nz zero bits, 00 00 ff ff) after its last slot. */
__device__ __forceinline__ void emit_run(const uint32_t *tab, const uint16_t *slots, uint32_t beg, uint32_t end, uint32_t start, uint32_t acc0,
                                         bool owns_first, bool trailer, uint32_t nz, uint32_t *slotw, uint64_t pkeep)
{
    uint64_t acc = owns_first ? (uint64_t)acc0 : 0ull;
    uint32_t nacc = start & 31, wpos = start >> 5;
    bool partial = !owns_first && nacc != 0;
#define QZ_EMIT_FLUSH() do { if (nacc >= 32) { if (partial) { atomicOr(slotw + wpos, (uint32_t)acc); partial = false; } else slot_st(slotw + wpos, (uint32_t)acc); \
                                               acc >>= 32; nacc -= 32; wpos++; } } while (0)
    uint4 qn = beg < end ? tok_ld4(reinterpret_cast<const uint32_t *>(slots + beg), pkeep) : make_uint4(0, 0, 0, 0);
    for (uint32_t j = beg; j < end; j += 8) {
        const uint32_t w[4] = { qn.x, qn.y, qn.z, qn.w };
        if (j + 8 < end) qn = tok_ld4(reinterpret_cast<const uint32_t *>(slots + j + 8), pkeep);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (j + k < end) {
                uint32_t val, nb;
                slot_code(tab, (w[k >> 1] >> ((k & 1) * 16)) & 0xffffu, val, nb);
                acc |= (uint64_t)val << nacc; nacc += nb;
                QZ_EMIT_FLUSH();
            }
        }
    }
    if (trailer) {          /* owner of the end-of-block slot: empty stored block */
        nacc += nz; QZ_EMIT_FLUSH();
        nacc += 16; QZ_EMIT_FLUSH();
        acc |= (uint64_t)0xffffu << nacc; nacc += 16; QZ_EMIT_FLUSH();
    }
    if ((uint32_t)acc) atomicOr(slotw + wpos, (uint32_t)acc);
#undef QZ_EMIT_FLUSH
}
/* slots per lane for a run of NT slots: a multiple of 8 */
__device__ __forceinline__ uint32_t run_length(uint32_t NT) { return (((NT + 31) >> 5) + 7) & ~7u; }

/* stored block for one piece: the piece starts byte-aligned, so the 3 header bits + pad are one byte.
 * The bytes come from global memory again (the shared window already belongs to somebody else);
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

/* one piece as its own block: everything after the histogram.  slots[0..ps.nslots) are the piece's slots; the end-of-block
 * slot is appended here. */
__device__ __forceinline__ void finish_piece(const QzbCompressJob &job, BlockCoder &bc, uint16_t *slots, const uint16_t *s_lentab, uint32_t lane, const PieceState &ps,
                                             uint32_t extra_total, uint64_t pkeep QZ_TARG)
{
    const uint32_t g = ps.g, n = ps.n;
    const bool bfinal = ps.bfinal;
    uint8_t *slot = job.slots + (size_t)g * job.slot_stride;
    uint32_t *slotw = reinterpret_cast<uint32_t *>(slot);
    if (lane == 0) tok16_st(slots + ps.nslots, 256, pkeep);
    __syncwarp();

    uint32_t out_bytes = 0;
    int btype = choose_block(bc.cs, bc.hist, extra_total, (5 + n) * 8, job.static_huffman, lane QZ_TPASS);
    if (n == 0) btype = 1;

    if (btype == 0) out_bytes = stored_piece(slot, ps.src, n, bfinal, lane);
    else {
        uint32_t hb, pend2;
        open_block(bc.cs, bc.hist, btype, bfinal, slotw, s_lentab, lane, &hb, &pend2 QZ_TPASS);
        /* Every lane codes a contiguous run of slots: pass 1 adds up the run's bit length, a warp scan turns the lengths
         * into bit offsets, pass 2 packs the run through a private 64-bit accumulator straight into the output.  Words that
         * hold a run boundary are zeroed first and receive their parts by atomic OR; every other word is written whole by
         * exactly one lane.  The end-of-block code is the last slot; the lane that owns it also appends the byte-alignment
         * trailer. */
        const uint32_t NT = ps.nslots + 1;
        const uint32_t R = run_length(NT);
        const uint32_t beg = min(lane * R, NT), end = min(beg + R, NT);
        const uint32_t mybits = count_run_bits(bc.hist, slots, beg, end, pkeep);
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
        emit_run(bc.hist, slots, beg, end, start, pend2, lane == 0, !bfinal && beg < NT && end == NT, nz, slotw, pkeep);
        out_bytes = (end_bit2 + 7) >> 3;
        __syncwarp();
    }
    if (lane == 0) job.piece_len[g] = out_bytes;
    __syncwarp();
    QZ_MARK(8);
}

/* take a free buffer (piece buffer or window unit): one bit per buffer in *busy (set = free); a warp that finds none sleeps
 * with exponential back-off instead of spinning on the issue slots the working warps need */
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

template <int PIECE_LOG2, int HB>
__global__ void __launch_bounds__(QZ_PIECES_MAX_WARPS * 32) qzb_deflate_pieces_kernel(QzbCompressJob job, int nbuf)
{
    constexpr int PIECE = 1 << PIECE_LOG2;
    static_assert(sizeof(WarpPriv<HB>) == (sizeof(uint16_t) << HB), "the block coder must fit in the hash table");
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
    uint16_t *slots = reinterpret_cast<uint16_t *>(job.tok_scratch + (size_t)gwarp * QZB_TOK_STRIDE(PIECE));
    const uint64_t pkeep = l2_policy_keep();

#ifdef QZ_PHASE_CLOCKS
    long long tlast = clock64();
#endif
    for (;;) {
        uint32_t g = 0;
        if (lane == 0) g = atomicAdd(job.ticket, 1u);
        g = __shfl_sync(FULL, g, 0);
        if (g >= job.npieces) break;
        const uint32_t b = take_buffer(&s_busy[0], lane);
        QZ_MARK(0);
        PieceState ps;
        piece_geometry<PIECE_LOG2>(job, g, ps);
        {
            uint8_t *piece = bufs[b].bytes + QZM_FRONT_PAD;
            const uint32_t c = load_and_checksum<PIECE>(piece, ps.src, ps.n, true, job.fmt, s_crc_tab, s_xstrip, lane);
            if (lane == 0) job.piece_crc[g] = c;
            for (uint32_t i = lane; i < (1u << HB) / 2; i += 32) reinterpret_cast<uint32_t *>(ws.u.table)[i] = 0xffffffffu;
            __syncwarp();
            QZ_MARK(1);
            QzmDeflateSink sink = { slots, 0, pkeep };
            qzm_match_piece<1>(piece, ps.n, 0, ps.n, ws.u.table, 1u << HB, sink, lane);
            ps.nslots = sink.nslots;
            QZ_MARK(2);
        }
        give_buffer(&s_busy[0], b, lane);
        BlockCoder &bc = ws.u.b;
        for (uint32_t i = lane; i < QZ_HIST_WORDS; i += 32) bc.hist[i] = 0;
        __syncwarp();
        const uint32_t extra_total = warp_sum(slot_hist(bc.hist, slots, ps.nslots, s_lentab, lane, pkeep));
        __syncwarp();
        QZ_MARK(3);
        finish_piece(job, bc, slots, s_lentab, lane, ps, extra_total, pkeep QZ_TPASS);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Window kernel (default for hw_buff_sz >= 64 KiB): a WINDOW is 64 KiB of one chunk and becomes ONE deflate block, what
 * the QAT engine emits for a 64 KiB request with its 32 KiB history (reference src/qatzip_utils.c:270-291).
 *
 * A CTA (one per SM, 32 warps) holds ONE window at a time in shared memory, next to thirty hash tables that take all the
 * shared memory there is (about 2600 entries each): thirty MATCHERS, each with a sub-piece of 2208 bytes (69 tiles; the
 * last one 1504), and two CODERS that take turns, window by window.  The coder of window k draws it and has it copied in
 * by one TMA bulk copy (cp.async.bulk, completion on an mbarrier) that is issued the moment the matchers have finished
 * with window k - 1 and runs while they emit window k - 2; it takes the window's checksum while the matchers match it.
 * The matchers run the stages of qz_match.cuh (prepass, seed, match: every position has the whole window in front of it
 * as history while the sub-pieces are matched concurrently), leave their tokens as slots in the L2 scratch and add them
 * to the window's histogram.  The same coder then turns that histogram into code tables and the block header (sort,
 * Huffman lengths, header plan, canonical codes: several thousand mostly serial instructions) WHILE the matchers are
 * already matching window k + 1 (fetched and checksummed by the other coder); when they are done with that, they meet
 * the coders, and emit window k: every warp counts its slots' bits, the totals are scanned through shared memory,
 * boundary words are zeroed, and every lane packs its run at its bit offset inside the window's output (the slot of its
 * first 8 KiB piece); the lane that codes the end-of-block slot appends the byte-aligning empty stored block unless the
 * block is final.  Slots, histograms and code tables are double-buffered by window parity.  Incompressible window ->
 * every 8 KiB piece a stored block in its own slot.  Warps meet at named barriers (matchers only / everybody).
 * The window's checksum goes to piece_crc[] of the window's first 8 KiB piece (the framing kernel combines per window). */
#define QZ_WINDOW 65536u
#define QZ_WINDOW_PIECES 8u             /* 8 KiB job pieces (slots, lengths) per window */
#define QZW_MATCHERS 30u
#define QZW_CODERS 2u
#define QZW_WARPS (QZW_MATCHERS + QZW_CODERS)
#define QZW_SUB 2208u                   /* bytes of a matcher's sub-piece: 69 tiles, a multiple of 16 */
static_assert(QZW_SUB % 32 == 0 && QZW_SUB * QZW_MATCHERS >= QZ_WINDOW && QZW_SUB * (QZW_MATCHERS - 1) < QZ_WINDOW, "sub-pieces tile the window");
static_assert(QZW_WARPS == QZ_GROUPS_MAX_WARPS, "the kernel's launch bound");
struct WindowShared {                   /* one per window parity */
    uint32_t ticket, btype, hb, pend;
    uint32_t nslots[QZW_WARPS], bits[QZW_WARPS], extra[QZW_WARPS];
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
/* bytes of one unit: the window with its pads, then the matchers' tables of `tent` entries (rounded up to 16 bytes each) */
__host__ __device__ __forceinline__ uint32_t window_table_stride(uint32_t tent) { return (tent + 8u) & ~7u; }       /* u16 entries */
__host__ __device__ __forceinline__ uint32_t window_unit_bytes(uint32_t tent) { return QZM_FRONT_PAD + QZ_WINDOW + QZM_TAIL_PAD + QZW_MATCHERS * 2u * window_table_stride(tent); }
/* 32-bit words of slot scratch per warp of the window kernel: two windows in flight */
#define QZW_TOK_WORDS (2u * QZB_TOK_STRIDE(QZW_SUB))

/* where window gi lies */
struct WindowGeom {
    uint32_t g0, wlen, nsub, npc; bool gfinal; const uint8_t *wsrc;
};
__device__ __forceinline__ WindowGeom window_geometry(const QzbCompressJob &job, uint32_t gi)
{
    WindowGeom w;
    const uint32_t wpc = job.pieces_per_chunk / QZ_WINDOW_PIECES;     /* windows per chunk */
    const uint32_t chunk = gi / wpc, blk = gi - chunk * wpc;
    w.g0 = chunk * job.pieces_per_chunk + blk * QZ_WINDOW_PIECES;
    const uint64_t chunk_off = (uint64_t)chunk * job.chunk_sz;
    const uint64_t rem = job.src_len > chunk_off ? job.src_len - chunk_off : 0;
    const uint32_t chunk_len = rem < job.chunk_sz ? (uint32_t)rem : job.chunk_sz;
    const uint32_t win_off = blk * QZ_WINDOW;
    w.wlen = min(QZ_WINDOW, chunk_len - win_off);                     /* > 0: only windows with data are counted */
    w.nsub = (w.wlen + QZW_SUB - 1) / QZW_SUB;                        /* sub-pieces with data */
    w.npc = (w.wlen + 8191u) >> 13;                                   /* 8 KiB job pieces with data */
    w.gfinal = (win_off + w.wlen == chunk_len) && (job.fmt != QZB_FMT_RAW || (chunk == job.nchunks - 1 && job.last));
    w.wsrc = job.src + chunk_off + win_off;
    return w;
}

/* Checksum of the window in shared memory by one warp (a coder): CRC-32, or packed Adler-32 sums for zlib streams.
 * Lane i owns the strip [m - (32 - i) S, m - (31 - i) S) of the window's whole words (m = n & ~3) with S = 2052 bytes (513
 * words: the lanes' loads fall into different banks), as three thirds of 684 bytes with independent dependency chains,
 * four bytes per step through four tables (slicing by 4).  The strips' remainders are "pure" (no initial value, no final
 * inversion): positions in front of the window count as zero bytes and leave a pure remainder untouched, so short windows
 * take the same code; the up to three bytes behind m continue the joined remainder; crc32(M) = pure(M) ^ crc32(|M| zero bytes).
 * s_xw[0] = x^(8 * 684), s_xw[1 + k] = x^(8 * 2052 * 2^k), s_xw[6] = crc32 of 65536 zero bytes. */
#define QZW_CK_STRIP 2052
#define QZW_CK_THIRD 684
__device__ __noinline__ uint32_t window_checksum(const uint8_t *win, uint32_t n, int fmt, const uint32_t *s_crc_tab /* [4][256] */, const uint32_t *s_xw, uint32_t lane)
{
    if (fmt == QZB_FMT_ZLIB) {
        /* Adler-32 sums of the strip (2052 < NMAX: no reduction inside), joined up a tree like the CRC terms */
        const int hi = (int)n - (int)((31 - lane) * QZW_CK_STRIP), lo = hi - QZW_CK_STRIP;
        uint32_t s1 = 0, s2 = 0;
        for (int i = lo < 0 ? 0 : lo; i < hi; i++) { s1 += win[i]; s2 += s1; }
        s2 %= QZ_ADLER_P;
        uint64_t len = (uint64_t)(hi > 0 ? hi - (lo < 0 ? 0 : lo) : 0);
#pragma unroll 1
        for (int lv = 0; lv < 5; lv++) {
            const uint32_t o1 = __shfl_down_sync(FULL, s1, 1u << lv), o2 = __shfl_down_sync(FULL, s2, 1u << lv);
            const uint64_t olen = __shfl_down_sync(FULL, len, 1u << lv);
            if ((lane & ((2u << lv) - 1)) == 0) { qz_adler_join(&s1, &s2, o1, o2, olen); len += olen; }
        }
        return __shfl_sync(FULL, qz_adler_pack(s1, s2), 0);
    }
    const uint32_t m = n & ~3u;
    const int lo = (int)m - (int)((32 - lane) * QZW_CK_STRIP);
    const uint32_t *T0 = s_crc_tab, *T1 = s_crc_tab + 256, *T2 = s_crc_tab + 512, *T3 = s_crc_tab + 768;
    uint32_t c0 = 0, c1 = 0, c2 = 0;
#define QZ_CRC_STEP4(c, w) do { c ^= (w); c = T3[c & 0xff] ^ T2[(c >> 8) & 0xff] ^ T1[(c >> 16) & 0xff] ^ T0[c >> 24]; } while (0)
#pragma unroll 1
    for (int j = 0; j < QZW_CK_THIRD; j += 4) {
        const int q0 = lo + j, q1 = q0 + QZW_CK_THIRD, q2 = q1 + QZW_CK_THIRD;
        const uint32_t w0 = q0 >= 0 ? *reinterpret_cast<const uint32_t *>(win + q0) : 0u;
        const uint32_t w1 = q1 >= 0 ? *reinterpret_cast<const uint32_t *>(win + q1) : 0u;
        const uint32_t w2 = q2 >= 0 ? *reinterpret_cast<const uint32_t *>(win + q2) : 0u;
        QZ_CRC_STEP4(c0, w0); QZ_CRC_STEP4(c1, w1); QZ_CRC_STEP4(c2, w2);
    }
#undef QZ_CRC_STEP4
    uint32_t c = qz_gf2_mul(qz_gf2_mul(c0, s_xw[0]) ^ c1, s_xw[0]) ^ c2;
#pragma unroll 1
    for (int lv = 0; lv < 5; lv++) {
        const uint32_t other = __shfl_down_sync(FULL, c, 1u << lv);   /* right neighbour block */
        if ((lane & ((2u << lv) - 1)) == 0) c = qz_gf2_mul(c, s_xw[1 + lv]) ^ other;
    }
    c = __shfl_sync(FULL, c, 0);
    for (uint32_t i = m; i < n; i++) c = T0[(c ^ win[i]) & 0xff] ^ (c >> 8);
    return c ^ (n == QZ_WINDOW ? s_xw[6] : ~qz_gf2_mul(0xffffffffu, qz_crc_xpow8(n)));
}

/* WAYS: entries of a hash bucket the match stage looks at: 1 (levels 1-5), 2 (levels 6 and up) */
template <int WAYS>
__global__ void __launch_bounds__(QZ_GROUPS_MAX_WARPS * 32) qzb_deflate_window_kernel(QzbCompressJob job)
{
    constexpr uint32_t SUB = QZW_SUB, NM = QZW_MATCHERS, NW = QZW_WARPS;
    constexpr int PIECE = 1 << 13;
    static_assert(sizeof(BlockCoder) % 16 == 0, "block coders are laid end to end");
    QZ_DYN_SMEM(smem_raw);
    __shared__ uint32_t s_crc_tab[4 * 256];         /* slicing by 4 */
    __shared__ uint32_t s_xw[7];
    __shared__ uint16_t s_lentab[256];
    __shared__ uint64_t s_mbar[1];          /* "window k is in shared memory", one phase per window */
    __shared__ WindowShared s_win[2];

    const uint32_t lane = lane_id(), wg = threadIdx.x >> 5;
    const uint32_t tent = job.tent, tstride = window_table_stride(tent);
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
        uint32_t e = qz_crc_table_entry(i);
        s_lentab[i] = len_table_entry(i);
        s_crc_tab[i] = e;
        for (int t = 1; t < 4; t++) { e = (e >> 8) ^ qz_crc_table_entry(e & 0xff); s_crc_tab[t * 256 + i] = e; }
    }
    if (threadIdx.x == 0) s_xw[0] = qz_crc_xpow8(QZW_CK_THIRD);
    if (threadIdx.x >= 1 && threadIdx.x < 6) s_xw[threadIdx.x] = qz_crc_xpow8((uint64_t)QZW_CK_STRIP << (threadIdx.x - 1));
    if (threadIdx.x == 6) s_xw[6] = ~qz_gf2_mul(0xffffffffu, qz_crc_xpow8(QZ_WINDOW));
    if (threadIdx.x == 0) qz_mbar_init(&s_mbar[0]);
    __syncthreads();

    constexpr uint32_t bar_m = 1, bar_f = 2;                           /* matchers only / everybody */
    BlockCoder *coders = reinterpret_cast<BlockCoder *>(smem_raw + window_unit_bytes(tent));
    uint8_t *win = smem_raw + QZM_FRONT_PAD;
    uint16_t *tables = reinterpret_cast<uint16_t *>(smem_raw + QZM_FRONT_PAD + QZ_WINDOW + QZM_TAIL_PAD);
    uint64_t *mbar = &s_mbar[0];
    const uint64_t pkeep = l2_policy_keep();
#ifdef QZ_PHASE_CLOCKS
    long long tlast = clock64();
#endif

    if (wg >= NM) {
        /* ---- the coders.  Coder c owns the windows k with k & 1 == c: it draws window k and has it copied in (TMA bulk
         * copy, completion on the mbarrier) once the matchers are done with window k - 1, checksums it while the matchers
         * match it, and builds its codes and block header while they match window k + 1. ---- */
        const uint32_t me = wg - NM;
        for (uint32_t k = 0;; k++) {
            WindowShared &G = s_win[k & 1];
            if ((k & 1) == me) {
                /* fetch window k (every matcher is done reading window k - 1) */
                uint32_t tk = 0;
                if (lane == 0) tk = atomicAdd(job.ticket, 1u);
                tk = __shfl_sync(FULL, tk, 0);
                uint32_t bulk = 0, wlen = 0, g0 = 0;
                const uint8_t *src = job.src;
                if (tk < job.ngroups) {
                    const WindowGeom w = window_geometry(job, tk);
                    src = w.wsrc; wlen = w.wlen; g0 = w.g0;
                    bulk = (reinterpret_cast<uintptr_t>(src) & 15) == 0 ? w.wlen & ~15u : 0u;       /* the rest by hand: a ragged tail, or everything from an unaligned source */
                    for (uint32_t i = bulk + lane; i < w.wlen; i += 32) win[i] = src[i];
                    for (uint32_t i = lane; i < QZM_TAIL_PAD; i += 32) win[w.wlen + i] = 0;
                }
                __syncwarp();
                if (lane == 0) { G.ticket = tk; qz_bulk_load_arrive(win, src, bulk, mbar); }
                __syncwarp();
                QZ_MARK(0);
                if (tk < job.ngroups) {
                    qz_mbar_wait(mbar, k);
                    const uint32_t ck = window_checksum(win, wlen, job.fmt, s_crc_tab, s_xw, lane);
                    if (lane == 0) job.piece_crc[g0] = ck;
                }
                QZ_MARK(1);
            } else if (k) {
                /* codes and block header of window k - 1 */
                WindowShared &P = s_win[(k - 1) & 1];
                BlockCoder &C = coders[(k - 1) & 1];
                const WindowGeom w = window_geometry(job, P.ticket);
                uint32_t extra_total = 0;
                for (uint32_t i = 0; i < NM; i++) extra_total += P.extra[i];
                const int btype = choose_block(C.cs, C.hist, extra_total, (5 * w.npc + w.wlen) * 8, job.static_huffman, lane QZ_TPASS);
                uint32_t hb = 0, pend = 0;
                if (btype) open_block(C.cs, C.hist, btype, w.gfinal, reinterpret_cast<uint32_t *>(job.slots + (size_t)w.g0 * job.slot_stride), s_lentab, lane, &hb, &pend QZ_TPASS);
                if (lane == 0) { P.btype = (uint32_t)btype; P.hb = hb; P.pend = pend; }
                QZ_MARK(11);
            }
            group_bar<NW * 32>(bar_f);
            QZ_MARK(10);                /* coders: waiting for the matchers */
            if (G.ticket >= job.ngroups) break;
        }
        return;
    }

    /* ---- the matchers ---- */
    const uint32_t gwarp = blockIdx.x * NM + wg;
    uint16_t *slots2 = reinterpret_cast<uint16_t *>(job.tok_scratch + (size_t)gwarp * QZW_TOK_WORDS);
    uint16_t *table = tables + (size_t)wg * tstride;
    const uint32_t p0 = wg * SUB;
    uint32_t prev_gi = 0;
    for (uint32_t k = 0;; k++) {
        const uint32_t b = k & 1;
        WindowShared &G = s_win[b];
        BlockCoder &C = coders[b];
        uint16_t *slots = slots2 + (size_t)b * 2 * QZB_TOK_STRIDE(SUB);
        qz_mbar_wait(mbar, k);
        QZ_MARK(0);
        const uint32_t gi = G.ticket;
        const bool have = gi < job.ngroups;
        if (have) {
            const WindowGeom w = window_geometry(job, gi);
            const uint32_t n = w.wlen > p0 ? min(SUB, w.wlen - p0) : 0u;                  /* bytes of this warp's sub-piece */
            const bool last_in_win = n != 0 && p0 + n == w.wlen;
            /* (a sub-piece's last three positions hash bytes of the next one: they are left to the match stage) */
            if (n) qzm_prepass<WAYS>(win, p0 + n, p0, p0 + n, table, tent, lane);
            group_bar<NM * 32>(bar_m);
            /* every matcher is past the emission of window k - 2: its code tables (same parity as window k) can go */
            if (wg == 0) { for (uint32_t i = lane; i < QZ_HIST_WORDS; i += 32) C.hist[i] = 0; }
            qzm_seed_tables<WAYS>(tables, tstride, w.nsub, tent, threadIdx.x, NM * 32);
            group_bar<NM * 32>(bar_m);
            QZ_MARK(15);
            QzmDeflateSink sink = { slots, 0, pkeep };
            if (n) qzm_match_piece<WAYS>(win, w.wlen, p0, p0 + n, table, tent, sink, lane);
            QZ_MARK(2);
            /* histogram of the warp's slots into the window's; the last sub-piece carries the end-of-block slot */
            const uint32_t extra = warp_sum(slot_hist(C.hist, slots, sink.nslots, s_lentab, lane, pkeep));
            if (lane == 0) {
                if (last_in_win) tok16_st(slots + sink.nslots, 256, pkeep);
                G.nslots[wg] = sink.nslots + (last_in_win ? 1u : 0u);
                G.extra[wg] = extra;
            }
            QZ_MARK(3);
        }
        group_bar<NW * 32>(bar_f);
        QZ_MARK(9);                 /* waiting for the slowest matcher and for the coders */
        /* ---- emission of window k - 1, whose code tables its coder has just finished ---- */
        if (k) {
            WindowShared &P = s_win[b ^ 1];
            const uint32_t *tab = coders[b ^ 1].hist;
            const uint16_t *pslots = slots2 + (size_t)(b ^ 1) * 2 * QZB_TOK_STRIDE(SUB);
            const WindowGeom w = window_geometry(job, prev_gi);      /* (P.ticket may already hold the ticket of window k + 1) */
            const uint32_t n = w.wlen > p0 ? min(SUB, w.wlen - p0) : 0u;
            const bool last_in_win = n != 0 && p0 + n == w.wlen;
            uint32_t *slotw = reinterpret_cast<uint32_t *>(job.slots + (size_t)w.g0 * job.slot_stride);
            if (P.btype == 0) {
                /* incompressible window: every 8 KiB piece is its own stored block in its own slot */
                if (wg < w.npc) {
                    const uint32_t pn = min((uint32_t)PIECE, w.wlen - wg * PIECE);
                    const uint32_t out_bytes = stored_piece(job.slots + (size_t)(w.g0 + wg) * job.slot_stride, w.wsrc + wg * PIECE, pn, w.gfinal && wg == w.npc - 1, lane);
                    if (lane == 0) job.piece_len[w.g0 + wg] = out_bytes;
                }
            } else {
                const uint32_t NT = P.nslots[wg];
                const uint32_t R = run_length(NT);
                const uint32_t beg = min(lane * R, NT), end = min(beg + R, NT);
                const uint32_t mybits = count_run_bits(tab, pslots, beg, end, pkeep);
                uint32_t incl = mybits;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(FULL, incl, o); if (lane >= (uint32_t)o) incl += y; }
                if (lane == 31) P.bits[wg] = incl;
                QZ_MARK(12);            /* count pass */
                group_bar<NM * 32>(bar_m);
                QZ_MARK(13);
                uint32_t before = P.hb, total = P.hb;
#pragma unroll
                for (uint32_t i = 0; i < NM; i++) {
                    const uint32_t bi = P.bits[i];
                    if (i < wg) before += bi;
                    total += bi;
                }
                const uint32_t end_bit = total;
                const uint32_t nz = 3 + ((0u - (end_bit + 3)) & 7);
                const uint32_t end_bit2 = w.gfinal ? end_bit : end_bit + nz + 32;
                const uint32_t start = before + incl - mybits;
                slotw[start >> 5] = 0;
                if (wg == NM - 1 && lane == 31) slotw[end_bit2 >> 5] = 0;
                group_bar<NM * 32>(bar_m);
                QZ_MARK(14);
                /* the lane that codes the end-of-block slot appends the trailer; it is the last slot of the window */
                const bool owns_eob = last_in_win && beg < NT && end == NT;
                emit_run(tab, pslots, beg, end, start, P.pend, wg == 0 && lane == 0, !w.gfinal && owns_eob, nz, slotw, pkeep);
                if (lane == 0 && wg < w.npc) job.piece_len[w.g0 + wg] = wg == 0 ? (end_bit2 + 7) >> 3 : 0u;        /* pieces behind a ragged end have no entry */
            }
            __syncwarp();
            QZ_MARK(8);
        }
        if (!have) break;
        prev_gi = gi;
    }
}

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
    /* checksum units: pieces, or -- window kernels -- 64 KiB windows whose checksum sits at their first piece */
    const uint32_t CKU = job.ngroups ? 65536u : PIECE, CKS = job.ngroups ? 8u : 1u;
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
                for (uint32_t k = 0; k * CKU < chunk_len; k++) {
                    const uint32_t pn = min(CKU, chunk_len - k * CKU), w = job.piece_crc[g0 + k * CKS];
                    qz_adler_join(&s1, &s2, w & 0xffffu, w >> 16, pn);
                }
                ck = qz_adler_finish(s1, s2, chunk_len);
            }
            ck = __shfl_sync(FULL, ck, 0);
        } else {
            ck = 0;
            for (uint32_t k = lane; k * CKU < chunk_len; k += 32) {
                const uint32_t after = chunk_len - min(chunk_len, (k + 1) * CKU);      /* input bytes behind checksum unit k */
                ck ^= qz_gf2_mul(job.piece_crc[g0 + k * CKS], qz_crc_xpow8(after));
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

extern "C" int qzb_deflate_max_warps(int window) { return window ? QZ_GROUPS_MAX_WARPS : QZ_PIECES_MAX_WARPS; }

/* shared memory of the window kernel: the window, thirty tables of `tent` entries, two block coders */
extern "C" size_t qzb_deflate_window_smem_bytes(int tent) { return (size_t)window_unit_bytes((uint32_t)tent) + 2 * sizeof(BlockCoder); }
/* the most entries a table can have: what the 227 KB per CTA leave after the kernel's static shared memory, the window and the
 * block coders */
extern "C" int qzb_deflate_window_max_tent(void)
{
    cudaFuncAttributes a;
    if (cudaFuncGetAttributes(&a, qzb_deflate_window_kernel<1>) != cudaSuccess) { (void)cudaGetLastError(); return 256; }
    const size_t cap = 227 * 1024 - a.sharedSizeBytes - 64;
    int tent = 256;
    while (tent + 8 <= 32768 && qzb_deflate_window_smem_bytes(tent + 8) <= cap) tent += 8;
    return tent;
}
/* 32-bit words of slot scratch the window kernel needs for a grid of CTAs */
extern "C" size_t qzb_deflate_window_tok_words(int grid) { return (size_t)grid * QZW_MATCHERS * QZW_TOK_WORDS; }

/* window kernel (one deflate block per 64 KiB window), one CTA of 32 warps per SM; job->ngroups and job->tent set;
 * ways: 1, or 2 for the deeper search of compression levels 6 and up */
template <int WAYS>
static cudaError_t launch_window(const QzbCompressJob &job, int grid, cudaStream_t st)
{
    const size_t smem = qzb_deflate_window_smem_bytes((int)job.tent);
    cudaError_t e = cudaFuncSetAttribute(qzb_deflate_window_kernel<WAYS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    qzb_deflate_window_kernel<WAYS><<<grid, QZW_WARPS * 32, smem, st>>>(job);
    return cudaGetLastError();
}
extern "C" cudaError_t qzb_launch_deflate_window(const QzbCompressJob *job, int grid, int ways, cudaStream_t st)
{
    if (job->pieces_per_chunk % QZ_WINDOW_PIECES || !job->ngroups || job->piece_log2 != 13 || job->tent < 256 || job->tent > 32768 || (job->tent & 1)) return cudaErrorInvalidValue;
    return ways == 2 ? launch_window<2>(*job, grid, st) : launch_window<1>(*job, grid, st);
}

extern "C" cudaError_t qzb_launch_frame(const QzbCompressJob *job, cudaStream_t st)
{
    qzb_chunk_sizes_kernel<<<(job->nchunks + 255) / 256, 256, 0, st>>>(*job);
    qzb_scan_kernel<<<1, 1024, 0, st>>>(job->chunk_total, job->chunk_off, job->nchunks);
    qzb_frame_kernel<<<job->nchunks, QZ_FRAME_WARPS * 32, 0, st>>>(*job);
    return cudaGetLastError();
}
#endif
