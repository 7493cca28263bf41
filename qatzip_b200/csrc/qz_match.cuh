/* qz_match.cuh -- the LZ77 match stage shared by the deflate and LZ4 compressors.
 *
 * What the QAT engine's history buffer does for a request (reference src/qatzip_utils.c:270-298: 32 KiB deflate window,
 * 64 KiB LZ4 blocks) is done here on a WINDOW: up to 64 KiB of one chunk sitting whole in shared memory, matched by up
 * to eight warps at once, each on its own PIECE (8 KiB) of it, every warp seeing everything in front of its position:
 *
 *   prepass   every warp records, in its own table of `tent` u16 entries, the last position of every 4-byte hash inside
 *             its piece                                                                                   (qzm_prepass)
 *   seed      a scan over the pieces, entry by entry: table k receives the most recent position of every hash in pieces
 *             0..k-1 -- what a sequential compressor's hash table would hold on entering piece k         (qzm_seed_tables)
 *   match     32 positions per step: all look up, then all insert.  A position inside a byte run takes the position before
 *             it as its candidate (distance 1) and is not inserted: the table keeps the run's start, which matches a later
 *             run of the same byte from ITS start; every other position's candidate is the table entry.  Candidates are
 *             verified and extended 12 bytes in the lane; the greedy parse of the tile -- which lanes start a token -- is
 *             found by pointer doubling over next(lane) = lane + max(1, L) instead of a serial walk, matches that reach
 *             the 12-byte cap are finished by the whole warp, a selected match takes over the literal in front of it when
 *             the byte before position and candidate agrees.                                      (qzm_match_piece)
 *
 * With one piece and an empty table the same routine is the private-window matcher of the per-piece kernels.
 * Tokens leave through a sink: 16-bit slots for deflate, (position, length, distance) records for LZ4.
 */
#ifndef QZ_MATCH_CUH
#define QZ_MATCH_CUH
#include <stdint.h>
#include <type_traits>
#include "qz_warp.cuh"

#define QZM_FULL 0xffffffffu
#define QZM_NONE 0xffffu
#define QZM_LANE_CAP 12u        /* bytes verified inside the lane; longer matches are finished by the warp */
#define QZM_FRONT_PAD 16u       /* readable bytes in front of a window (position - 1, position - 4 of its first positions) */
#define QZM_TAIL_PAD 48u        /* zero bytes behind the window's data */

__device__ __forceinline__ uint32_t qzm_hash(uint32_t v, uint32_t tent) { return __umulhi(v * 2654435761u, tent); }
__device__ __forceinline__ uint32_t qzm_ld32u(const uint32_t *w, uint32_t off) { return __funnelshift_r(w[off >> 2], w[(off >> 2) + 1], (off & 3) * 8); }

/* a position inside a byte run (the byte before it and its four bytes are all equal) is not recorded in the tables */
__device__ __forceinline__ bool qzm_run_interior(uint32_t v, uint32_t prevb, uint32_t p) { return p != 0 && v == prevb * 0x01010101u; }

/* WAYS = 1: table[h] = last position in [p0, p1) with hash h (QZM_NONE elsewhere).  WAYS = 2 (the deeper search of the higher
 * compression levels): the table is tent / 2 buckets of two entries, low half = the last position, high half = the one before.
 * `win` is the window's data (4-byte aligned, p0 a multiple of 32), n its length */
template <int WAYS>
__device__ __forceinline__ void qzm_prepass(const uint8_t *win, uint32_t n, uint32_t p0, uint32_t p1, uint16_t *table, uint32_t tent, uint32_t lane)
{
    for (uint32_t i = lane; i < (tent + 1) / 2; i += 32) reinterpret_cast<uint32_t *>(table)[i] = 0xffffffffu;
    __syncwarp();
    const uint32_t *ww = reinterpret_cast<const uint32_t *>(win);
    uint32_t *t32 = reinterpret_cast<uint32_t *>(table);
    const uint32_t sh = (lane & 3) * 8;
#pragma unroll 4
    for (uint32_t base = p0; base < p1; base += 32) {
        const uint32_t p = base + lane;
        const uint32_t *pw = ww + (p >> 2);
        const uint32_t v = __funnelshift_r(pw[0], pw[1], sh);
        /* ascending tiles: a later position overwrites an earlier one.  Equal hashes inside one tile are a write/write race
         * the hardware settles for one of the lanes -- positions less than 32 bytes apart, either will do. */
        const bool ins = p + 4 <= n && !qzm_run_interior(v, win[(int)p - 1], p);
        if (WAYS == 1) { if (ins) table[qzm_hash(v, tent)] = (uint16_t)p; }
        else {
            const uint32_t h = qzm_hash(v, tent / 2);
            const uint32_t w = ins ? t32[h] : 0u;
            __syncwarp();
            if (ins) t32[h] = (w << 16) | p;
            __syncwarp();
        }
    }
    __syncwarp();
}

/* tables[k] (k < npieces, `stride` u16 apart) <- the most recent position(s) of every hash in pieces 0..k-1.  Called by all
 * `nthreads` threads of the group between two group barriers; tables hold the prepass result on entry. */
template <int WAYS>
__device__ __forceinline__ void qzm_seed_tables(uint16_t *tables, uint32_t stride, uint32_t npieces, uint32_t tent, uint32_t tid, uint32_t nthreads)
{
    const uint32_t nw = (tent + 1) / 2, sw = stride / 2;
    uint32_t *t32 = reinterpret_cast<uint32_t *>(tables);
    for (uint32_t i = tid; i < nw; i += nthreads) {
        uint32_t carry = 0xffffffffu;
        for (uint32_t k = 0; k < npieces; k++) {
            const uint32_t w = t32[k * sw + i];
            t32[k * sw + i] = carry;
            if (WAYS == 1) {
                /* two independent entries per word */
                const uint32_t lo = (w & 0xffffu) != 0xffffu ? w & 0xffffu : carry & 0xffffu;
                const uint32_t hi = (w >> 16) != 0xffffu ? w & 0xffff0000u : carry & 0xffff0000u;
                carry = lo | hi;
            } else {
                /* one bucket per word: the piece's two most recent positions, filled up from what was carried */
                if ((w & 0xffffu) != 0xffffu) carry = (w >> 16) != 0xffffu ? w : (carry << 16) | (w & 0xffffu);
            }
        }
    }
}

/* ---- token sinks ---- */
/* deflate: 16-bit slots in an L2-resident scratch.  literal -> its byte; match -> 0x4000 | (len - 3), then 0x8000 | (dist - 1).
 * Every slot is coded on its own (a length slot and a distance slot each carry one code and its extra bits), so histogram,
 * bit count and emission run over slots, not over tokens. */
#define QZ_SLOT_LEN 0x4000u
#define QZ_SLOT_DIST 0x8000u
struct QzmDeflateSink {
    uint16_t *slots; uint32_t nslots; uint64_t pol;
    static constexpr bool kLz4 = false;
    static constexpr uint32_t kMinMatch = 4, kMaxMatch = 258, kMaxDist = 32768;
    __device__ __forceinline__ void put(uint32_t tokmask, uint32_t matchmask, uint32_t lane, uint32_t lt, uint32_t p, uint32_t v, uint32_t L, uint32_t dist)
    {
        if ((tokmask >> lane) & 1) {
            uint16_t *o = slots + nslots + __popc(tokmask & lt) + __popc(matchmask & lt);
            if ((matchmask >> lane) & 1) { tok16_st(o, (uint16_t)(QZ_SLOT_LEN | (L - 3)), pol); tok16_st(o + 1, (uint16_t)(QZ_SLOT_DIST | (dist - 1)), pol); }
            else tok16_st(o, (uint16_t)(v & 0xff), pol);
        }
        nslots += __popc(tokmask) + __popc(matchmask);
    }
};
/* LZ4: one record of two words per match: position | length << 16, distance (positions and lengths fit 16 bits: a window is at
 * most 64 KiB and a match ends inside its piece) */
struct QzmLz4Sink {
    uint32_t *recs; uint32_t nrec;
    static constexpr bool kLz4 = true;
    static constexpr uint32_t kMinMatch = 4, kMaxMatch = 65535, kMaxDist = 65535;
    __device__ __forceinline__ void put(uint32_t, uint32_t matchmask, uint32_t lane, uint32_t lt, uint32_t p, uint32_t, uint32_t L, uint32_t dist)
    {
        if ((matchmask >> lane) & 1) {
            uint32_t *o = recs + 2 * (nrec + __popc(matchmask & lt));
            o[0] = p | (L << 16); o[1] = dist;
        }
        nrec += __popc(matchmask);
    }
};

/* Greedy LZ77 parse of [p0, p1) of the window `win` (n bytes of data, pads as above).  `table` is seeded (or cleared, for a
 * private window); p0 is a multiple of 32.  LZ4 sinks: matches start at or before n - 12 and end at or before n - 5
 * (block end rules), and never inside the last piece's tail. */
template <int WAYS, class Sink>
__device__ __forceinline__ void qzm_match_piece(const uint8_t *win, uint32_t n, uint32_t p0, uint32_t p1, uint16_t *table, uint32_t tent, Sink &sink, uint32_t lane)
{
    const uint32_t *ww = reinterpret_cast<const uint32_t *>(win);
    uint32_t *t32 = reinterpret_cast<uint32_t *>(table);
    const uint32_t sh = (lane & 3) * 8, lt = qz_lanemask_lt();
    const uint32_t mstart_lim = Sink::kLz4 ? (n >= 13 ? n - 12 : 0u) : n;        /* LZ4: last position a match may start at */
    const uint32_t mend = Sink::kLz4 ? min(p1, n >= 5 ? n - 5 : 0u) : p1;         /* matches end at or before */
    uint32_t entry = 0;                 /* first position of the tile not covered by a match running in from the left */
    /* One tile.  EDGE = false is the copy for tiles well inside the sub-piece, where every bounds test is known to pass:
     * all 32 positions exist and have four bytes, a 12-byte match fits, (LZ4) matches may start. */
    auto tile = [&](const uint32_t base, auto edge_tag) {
        constexpr bool EDGE = decltype(edge_tag)::value;
        const uint32_t p = base + lane;
        const uint32_t *pw = ww + (p >> 2);
        const uint32_t w0 = pw[0], w1 = pw[1], w2 = pw[2], w3 = pw[3];
        const uint32_t v = __funnelshift_r(w0, w1, sh), v1 = __funnelshift_r(w1, w2, sh), v2 = __funnelshift_r(w2, w3, sh);
        const uint32_t prevb = win[(int)p - 1];
        const bool can = EDGE ? (p < p1 && p + 4 <= n) : true;
        const bool interior = qzm_run_interior(v, prevb, p);
        uint32_t t, t2 = QZM_NONE;
        if (WAYS == 1) {
            const uint32_t h = qzm_hash(v, tent);
            t = can ? table[h] : QZM_NONE;
            __syncwarp();
            if (can && !interior) table[h] = (uint16_t)p;       /* equal hashes inside the tile: see qzm_prepass */
            __syncwarp();
        } else {
            const uint32_t h = qzm_hash(v, tent / 2);
            const uint32_t w = can ? t32[h] : 0xffffffffu;
            __syncwarp();
            if (can && !interior) t32[h] = (w << 16) | p;
            __syncwarp();
            t = w & 0xffffu; t2 = w >> 16;
        }
        /* inside a byte run the candidate is the position before (distance 1); everywhere else the table's */
        const uint32_t cand1 = interior ? p - 1 : t;
        /* candidate bytes are fetched unconditionally (position 0 when there is none): no divergent verify branches */
        const bool has1 = can && cand1 != QZM_NONE && p - cand1 <= Sink::kMaxDist && (!EDGE || p <= mstart_lim);
        uint32_t cand = cand1, L;
        bool has = has1;
        {
            const uint32_t c = has1 ? cand1 : 0u, csh = (c & 3) * 8;
            const uint32_t *cw = ww + (c >> 2);
            const uint32_t c0 = cw[0], c1 = cw[1], c2 = cw[2], c3 = cw[3];
            const uint32_t x0 = __funnelshift_r(c0, c1, csh) ^ v, x1 = __funnelshift_r(c1, c2, csh) ^ v1, x2 = __funnelshift_r(c2, c3, csh) ^ v2;
            const uint32_t xx = x1 ? x1 : x2;
            L = (x1 ? 4u : 8u) + (xx ? (uint32_t)(__ffs(xx) - 1) >> 3 : 4u);     /* 4..12 = QZM_LANE_CAP */
            if (!has1 || x0) L = 0;
        }
        if (WAYS == 2) {
            /* the bucket's older entry: taken when it verifies longer (both reach the cap: the nearer one stays) */
            const bool has2 = can && !interior && t2 != QZM_NONE && p - t2 <= Sink::kMaxDist && (!EDGE || p <= mstart_lim);
            const uint32_t c = has2 ? t2 : 0u, csh = (c & 3) * 8;
            const uint32_t *cw = ww + (c >> 2);
            const uint32_t c0 = cw[0], c1 = cw[1], c2 = cw[2], c3 = cw[3];
            const uint32_t x0 = __funnelshift_r(c0, c1, csh) ^ v, x1 = __funnelshift_r(c1, c2, csh) ^ v1, x2 = __funnelshift_r(c2, c3, csh) ^ v2;
            const uint32_t xx = x1 ? x1 : x2;
            uint32_t L2 = (x1 ? 4u : 8u) + (xx ? (uint32_t)(__ffs(xx) - 1) >> 3 : 4u);
            if (!has2 || x0) L2 = 0;
            if (L2 > L) { L = L2; cand = t2; has = true; }
        }
        const uint32_t c = has ? cand : 0u;
        const uint32_t room = EDGE ? (mend > p ? mend - p : 0u) : 0xffffu;          /* bytes a match starting here may cover */
        if (EDGE) L = min(L, min(Sink::kMaxMatch, room));
        if (L < Sink::kMinMatch) L = 0;
        /* the byte in front of position and candidate agrees (and the candidate is not the window's first byte) */
        const bool back = L != 0 && c != 0 && prevb == win[(int)c - 1];

        /* greedy parse by pointer doubling: R = lanes visited from this lane on, N = the lane that walk has reached (a lane
         * whose token leaves the tile points at itself), E = where this lane's token ends */
        const uint32_t E = lane + (L ? L : 1u);
        uint32_t R = 1u << lane, N = E < 32 ? E : lane;
#pragma unroll
        for (int r = 0; r < 5; r++) { R |= __shfl_sync(QZM_FULL, R, N); N = __shfl_sync(QZM_FULL, N, N); }
        const uint32_t longmask = __ballot_sync(QZM_FULL, L >= QZM_LANE_CAP && (!EDGE || L < min(Sink::kMaxMatch, room)));
        uint32_t tokmask = 0, cur = entry, leave;
        for (;;) {
            const uint32_t Rc = __shfl_sync(QZM_FULL, R, cur);
            const uint32_t U = Rc & longmask;
            if (!U) { tokmask |= Rc; leave = __shfl_sync(QZM_FULL, E, 31 - __clz(Rc)); break; }
            /* the walk is right up to its first match that reached the lane cap: the warp finishes that one */
            const uint32_t m = __ffs(U) - 1;
            tokmask |= Rc & (QZM_FULL >> (31 - m));
            const uint32_t pm = base + m, cm = __shfl_sync(QZM_FULL, cand, m);
            const uint32_t mx = min(Sink::kMaxMatch, mend - pm);
            uint32_t Lm = QZM_LANE_CAP;
            while (Lm < mx) {
                const uint32_t kk = Lm + lane;
                const bool eq = kk < mx && win[pm + kk] == win[cm + kk];
                const uint32_t bal = __ballot_sync(QZM_FULL, eq);
                if (bal == QZM_FULL) { Lm += 32; continue; }
                Lm += __ffs(~bal) - 1; break;
            }
            Lm = min(Lm, mx);
            if (lane == m) L = Lm;
            cur = m + Lm;
            if (cur >= 32) { leave = cur; break; }
        }
        if (EDGE) tokmask &= __ballot_sync(QZM_FULL, p < p1);
        uint32_t matchmask = tokmask & __ballot_sync(QZM_FULL, L != 0);
        /* a selected match takes over the literal right in front of it when that byte agrees too */
        {
            const uint32_t lit = tokmask & ~matchmask;
            const bool ext = back && lane != 0 && ((matchmask >> lane) & (lit >> (lane - 1)) & 1) && L < Sink::kMaxMatch;
            tokmask &= ~(__ballot_sync(QZM_FULL, ext) >> 1);
            sink.put(tokmask, matchmask, lane, lt, p - (ext ? 1u : 0u), v, L + (ext ? 1u : 0u), p - cand);
        }
        entry = leave - 32;
    };
    /* tiles at or below this base are interior: no match of theirs can reach a limit (one compare per tile instead of three) */
    const int interior = min(min((int)mend - 32 - (int)QZM_LANE_CAP, (int)n - 35), (int)mstart_lim - 31);
    for (uint32_t base = p0; base < p1; base += 32) {
        if (entry >= 32) { entry -= 32; continue; }      /* tile lies inside a running match: nothing to code, not indexed */
        if ((int)base <= interior) tile(base, std::false_type());
        else tile(base, std::true_type());
    }
    __syncwarp();
}
#endif
