/* qz_deflate_split.cuh -- EXPERIMENTAL, compiled only with -DQZ_SPLIT_KERNEL (make ab ABFLAGS=-DQZ_SPLIT_KERNEL ABNAME=split)
 * and launched only with QZB200_GROUP=2.  Not part of the default build; validated on the SIMT emulator
 * (tests/test_emu_kernels.py::test_split_*), not yet on a GPU.
 *
 * The group kernel's warps spend a third of their time waiting while the piece buffers -- what shared
 * memory really limits -- are busy only as long as some warp happens to be matching.  Here the roles are split:
 *
 *   matcher warps  each own a piece buffer and a hash table for the whole launch and do nothing but phases 1-2 (load,
 *                  checksum, match, raw tokens to L2), one piece after the other, for the blocks the CTA has open;
 *   coder teams    four warps each, take a block whose eight pieces are all matched and run token pass -> codes -> count
 *                  -> emit on it (two pieces per warp), exactly the block format of the group kernel.
 *
 * A CTA keeps QZS_SLOTS blocks in flight.  A slot goes free -> opening -> filling -> ready -> coding -> free; matchers hand
 * themselves pieces of filling slots by atomic counter and open a new slot (one global ticket = one block) when none has
 * pieces left; the warp that finishes a block's last piece marks it ready.  All of this lives in shared memory; the only
 * traffic between the roles besides it is the token arrays in L2 (written and read with .cg accesses).
 */
#ifndef QZ_DEFLATE_SPLIT_CUH
#define QZ_DEFLATE_SPLIT_CUH

#define QZS_TEAM 4                  /* coder warps per team */
#define QZS_PPW (QZ_GROUP / QZS_TEAM)
#define QZS_SLOTS 4                 /* blocks in flight per CTA */
#define QZS_NONE 0xffffffffu
#define QZS_EXIT 0xfffffffeu
enum { QZS_FREE = 0, QZS_OPENING = 1, QZS_FILLING = 2, QZS_READY = 3, QZS_CODING = 4 };

struct SplitSlot {
    uint32_t state, gi, npieces, done, bfinal;
    uint32_t next;                  /* life << 8 | next piece to hand out: a claim is a CAS on the whole word, so a matcher that looked
                                     * at the slot in an earlier life can never take a piece of the current one by accident */
    uint32_t ntok[QZ_GROUP], nbytes[QZ_GROUP];
};
struct SplitTeam {
    uint32_t slot, btype, hb, pend;
    uint32_t ntok[QZ_GROUP];        /* tokens per piece including the end-of-block token of the last one */
    uint32_t bits[QZ_GROUP], extra[QZS_TEAM];
};
struct SplitShared {
    SplitSlot slot[QZS_SLOTS];
    uint32_t matchers_alive, no_more;
};

__device__ __forceinline__ uint32_t qzs_ld(const uint32_t *p) { return *reinterpret_cast<const volatile uint32_t *>(p); }

/* pieces of block gi that exist (a ragged last chunk ends early) */
__device__ __forceinline__ uint32_t qzs_block_pieces(const QzbCompressJob &job, uint32_t gi, uint32_t piece_log2)
{
    const uint32_t gpc = job.pieces_per_chunk / QZ_GROUP, chunk = gi / gpc, blk = gi - chunk * gpc;
    const uint64_t chunk_off = (uint64_t)chunk * job.chunk_sz;
    const uint64_t rem = job.src_len > chunk_off ? job.src_len - chunk_off : 0;
    const uint32_t chunk_len = rem < job.chunk_sz ? (uint32_t)rem : job.chunk_sz;
    const uint32_t first = (blk * QZ_GROUP) << piece_log2;
    const uint32_t left = chunk_len > first ? chunk_len - first : 0;
    const uint32_t np = (left + (1u << piece_log2) - 1) >> piece_log2;
    return np > QZ_GROUP ? (uint32_t)QZ_GROUP : np;
}

template <int PIECE_LOG2, int HB>
__global__ void __launch_bounds__(1024) qzb_deflate_split_kernel(QzbCompressJob job, int nmatch, int nteams)
{
    constexpr int PIECE = 1 << PIECE_LOG2;
    constexpr uint32_t STRIP = PIECE / 32 + 4;
    constexpr uint32_t TABLE_BYTES = 2u << HB, HIST_BYTES = (QZ_HIST_WORDS * 4 + 15) & ~15u;
    QZ_DYN_SMEM(smem_raw);
    __shared__ uint32_t s_crc_tab[256];
    __shared__ uint32_t s_xstrip[5];
    __shared__ uint16_t s_lentab[256];
    __shared__ SplitShared sh;
    __shared__ SplitTeam s_team[8];

    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) { s_crc_tab[i] = qz_crc_table_entry(i); s_lentab[i] = len_table_entry(i); }
    if (threadIdx.x < 5) s_xstrip[threadIdx.x] = qz_crc_xpow8((uint64_t)STRIP << threadIdx.x);
    if (threadIdx.x < QZS_SLOTS) { sh.slot[threadIdx.x].state = QZS_FREE; sh.slot[threadIdx.x].next = 0; sh.slot[threadIdx.x].npieces = 0; }
    if (threadIdx.x == 0) { sh.matchers_alive = (uint32_t)nmatch; sh.no_more = 0; }
    __syncthreads();

    /* dynamic shared memory: [piece buffer + hash table] per matcher, [histogram] per coder warp, [code scratch] per team */
    uint8_t *m_base = smem_raw;
    uint8_t *c_base = m_base + (size_t)nmatch * (sizeof(PieceBuf<PIECE_LOG2>) + TABLE_BYTES);
    uint8_t *t_base = c_base + (size_t)nteams * QZS_TEAM * HIST_BYTES;
    const uint64_t pkeep = l2_policy_keep();
    const uint32_t gpc = job.pieces_per_chunk / QZ_GROUP;
    uint32_t *tok_cta = job.tok_scratch + (size_t)blockIdx.x * QZS_SLOTS * QZ_GROUP * QZB_TOK_STRIDE(PIECE);
#ifdef QZ_PHASE_CLOCKS
    long long tlast = clock64();
#endif

    if (warp < (uint32_t)nmatch) {
        /* ------------------------------------------------------------------ matcher */
        uint8_t *piece = m_base + (size_t)warp * (sizeof(PieceBuf<PIECE_LOG2>) + TABLE_BYTES);
        uint16_t *table = reinterpret_cast<uint16_t *>(piece + sizeof(PieceBuf<PIECE_LOG2>));
        for (;;) {
            uint32_t s = QZS_NONE, k = 0;
            if (lane == 0) {
                uint32_t ns = 64;
                for (;;) {
                    bool opening = false;
                    for (uint32_t i = 0; i < QZS_SLOTS && s == QZS_NONE; i++) {
                        SplitSlot &S = sh.slot[i];
                        const uint32_t st = qzs_ld(&S.state);
                        if (st == QZS_OPENING) opening = true;
                        if (st == QZS_FILLING) {
                            const uint32_t w = qzs_ld(&S.next);                 /* written last when a slot is opened: the fields below belong to life w >> 8 or later */
                            if ((w & 0xffu) < qzs_ld(&S.npieces) && atomicCAS(&S.next, w, w + 1) == w) { s = i; k = w & 0xffu; }
                        }
                    }
                    if (s != QZS_NONE) break;
                    if (!qzs_ld(&sh.no_more)) {
                        /* nothing to hand out: open the next block in a free slot, if there is one */
                        for (uint32_t i = 0; i < QZS_SLOTS && s == QZS_NONE; i++) {
                            SplitSlot &S = sh.slot[i];
                            if (qzs_ld(&S.state) != QZS_FREE || atomicCAS(&S.state, (uint32_t)QZS_FREE, (uint32_t)QZS_OPENING) != QZS_FREE) continue;
                            const uint32_t gi = atomicAdd(job.ticket, 1u);
                            if (gi >= job.ngroups) {
                                sh.no_more = 1; __threadfence_block(); S.state = QZS_FREE;
                            } else {
                                /* new life: first close the hand-out word (no piece index is below 0xff), then the fields, then open it
                                 * with piece 0 taken by this warp -- a claim prepared against the previous life fails its CAS */
                                const uint32_t life = ((qzs_ld(&S.next) >> 8) + 1) << 8;
                                S.next = life | 0xffu;
                                __threadfence_block();
                                S.gi = gi; S.npieces = qzs_block_pieces(job, gi, PIECE_LOG2); S.done = 0; S.bfinal = 0;
                                for (int q = 0; q < QZ_GROUP; q++) { S.ntok[q] = 0; S.nbytes[q] = 0; }
                                __threadfence_block();
                                S.next = life | 1u;
                                __threadfence_block();
                                S.state = QZS_FILLING;
                                s = i; k = 0;
                            }
                            break;
                        }
                        if (s != QZS_NONE) break;
                    } else if (!opening) { s = QZS_EXIT; break; }       /* no block will get pieces any more */
                    __nanosleep(ns);
                    if (ns < 2048) ns <<= 1;
                }
            }
            s = __shfl_sync(FULL, s, 0); k = __shfl_sync(FULL, k, 0);
            if (s == QZS_EXIT) break;
            SplitSlot &S = sh.slot[s];
            const uint32_t gi = S.gi, chunk = gi / gpc, blk = gi - chunk * gpc;
            const uint32_t g = chunk * job.pieces_per_chunk + blk * QZ_GROUP + k;
            uint32_t *toks = tok_cta + ((size_t)s * QZ_GROUP + k) * QZB_TOK_STRIDE(PIECE);
            PieceState ps;
            QZ_MARK(0);
            phase12<PIECE_LOG2, HB>(job, piece, table, toks, s_crc_tab, s_xstrip, g, lane, ps QZ_TPASS);
            __syncwarp();
            if (lane == 0) {
                S.ntok[k] = ps.ntok; S.nbytes[k] = ps.n;
                if (ps.bfinal) atomicOr(&S.bfinal, 1u);
                __threadfence();                                   /* tokens (L2) and the lengths before the count */
                if (atomicAdd(&S.done, 1u) + 1 == S.npieces) { __threadfence_block(); S.state = QZS_READY; }
            }
            __syncwarp();
        }
        if (lane == 0) { __threadfence_block(); atomicSub(&sh.matchers_alive, 1u); }
        return;
    }

    /* ---------------------------------------------------------------------- coder team */
    const uint32_t cw = warp - (uint32_t)nmatch, team = cw / QZS_TEAM, tw = cw % QZS_TEAM, bar = 1 + team;
    if (team >= (uint32_t)nteams) return;
    SplitTeam &T = s_team[team];
    uint32_t *hist = reinterpret_cast<uint32_t *>(c_base + (size_t)cw * HIST_BYTES);
    GroupLead &L = *reinterpret_cast<GroupLead *>(t_base + (size_t)team * sizeof(GroupLead));
    for (;;) {
        if (tw == 0 && lane == 0) {
            uint32_t ns = 64, got = QZS_NONE;
            for (;;) {
                const uint32_t alive = qzs_ld(&sh.matchers_alive);
                for (uint32_t i = 0; i < QZS_SLOTS && got == QZS_NONE; i++)
                    if (qzs_ld(&sh.slot[i].state) == QZS_READY && atomicCAS(&sh.slot[i].state, (uint32_t)QZS_READY, (uint32_t)QZS_CODING) == QZS_READY) got = i;
                if (got != QZS_NONE) break;
                if (alive == 0) { got = QZS_EXIT; break; }         /* the matchers were gone before the scan: nothing can turn ready */
                __nanosleep(ns);
                if (ns < 2048) ns <<= 1;
            }
            __threadfence();
            T.slot = got;
        }
        group_bar<QZS_TEAM * 32>(bar);
        QZ_MARK(15);                /* coder: waiting for a block whose pieces are all matched */
        const uint32_t s = T.slot;
        if (s == QZS_EXIT) break;
        SplitSlot &S = sh.slot[s];
        const uint32_t gi = S.gi, chunk = gi / gpc, blk = gi - chunk * gpc;
        const uint32_t g0 = chunk * job.pieces_per_chunk + blk * QZ_GROUP;
        uint32_t *tok_slot = tok_cta + (size_t)s * QZ_GROUP * QZB_TOK_STRIDE(PIECE);
        PieceState ps[QZS_PPW];
        bool last_in_group[QZS_PPW];
#pragma unroll
        for (int j = 0; j < QZS_PPW; j++) {
            const uint32_t pi = tw * QZS_PPW + j;
            const uint64_t chunk_off = (uint64_t)chunk * job.chunk_sz;
            const uint64_t rem = job.src_len > chunk_off ? job.src_len - chunk_off : 0;
            const uint32_t chunk_len = rem < job.chunk_sz ? (uint32_t)rem : job.chunk_sz;
            const uint32_t p_off = (blk * QZ_GROUP + pi) << PIECE_LOG2;
            ps[j].g = g0 + pi; ps[j].n = S.nbytes[pi]; ps[j].ntok = S.ntok[pi]; ps[j].extra_total = 0;
            ps[j].src = job.src + chunk_off + p_off;
            const bool chunk_end = ps[j].n != 0 && p_off + ps[j].n == chunk_len;
            ps[j].bfinal = chunk_end && (job.fmt != QZB_FMT_RAW || (chunk == job.nchunks - 1 && job.last));
            last_in_group[j] = ps[j].n != 0 && (pi == QZ_GROUP - 1 || chunk_end);
        }
        const bool gfinal = S.bfinal != 0;
        uint32_t extra = 0;
#pragma unroll
        for (int j = 0; j < QZS_PPW; j++) {
            uint32_t *toks = tok_slot + (size_t)(tw * QZS_PPW + j) * QZB_TOK_STRIDE(PIECE);
            extra += token_pass(hist, toks, ps[j].ntok, s_lentab, lane, pkeep, j == 0);
            if (lane == 0 && last_in_group[j]) tok_st(toks + ps[j].ntok, 256u, pkeep);
        }
        extra = warp_sum(extra);
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < QZS_PPW; j++) T.ntok[tw * QZS_PPW + j] = ps[j].ntok + (last_in_group[j] ? 1u : 0u);
            T.extra[tw] = extra;
        }
        QZ_MARK(3);
        group_bar<QZS_TEAM * 32>(bar);
        QZ_MARK(9);                 /* waiting for the team's slowest token pass */
        /* mixed block: a block per piece (see the group kernel), the pieces take turns on the team's code scratch */
        bool mixed;
        {
            uint32_t hi = 0, lo = 0xffffffffu;
#pragma unroll
            for (int i = 0; i < QZ_GROUP; i++) {
                const uint32_t nb = S.nbytes[i];
                if (nb) { const uint32_t r = (T.ntok[i] << 10) / nb; hi = max(hi, r); lo = min(lo, r); }
            }
            mixed = hi > 920u && lo < 768u;
        }
        if (mixed) {
            for (uint32_t turn = 0; turn < QZ_GROUP; turn++) {
                if (turn / QZS_PPW == tw) {
                    const int j = (int)(turn % QZS_PPW);
                    if (ps[j].n) {
                        uint32_t *toks = tok_slot + (size_t)turn * QZB_TOK_STRIDE(PIECE);
                        if (lane == 0) tok_st(toks + ps[j].ntok, 256u, pkeep);
                        __syncwarp();
                        const uint32_t ex = hist_from_tokens(L.hist, toks, ps[j].ntok, lane, pkeep);
                        finish_piece(job, L.cs, L.hist, toks, lane, ps[j], ex, pkeep QZ_TPASS);
                    }
                }
                group_bar<QZS_TEAM * 32>(bar);
            }
        } else {
            if (tw == 0) {
                uint32_t extra_total = 0, nbytes = 0, npc = 0;
                for (int i = 0; i < QZS_TEAM; i++) extra_total += T.extra[i];
                for (int i = 0; i < QZ_GROUP; i++) { nbytes += S.nbytes[i]; npc += S.nbytes[i] ? 1u : 0u; }
                for (uint32_t q = lane; q < QZ_HIST_WORDS; q += 32) {
                    uint32_t f = 0;
#pragma unroll
                    for (int i = 0; i < QZS_TEAM; i++) f += reinterpret_cast<const uint32_t *>(c_base + (size_t)(team * QZS_TEAM + i) * HIST_BYTES)[q];
                    L.hist[q] = f;
                }
                __syncwarp();
                const int btype = choose_block(L.cs, L.hist, extra_total, (5 * npc + nbytes) * 8, job.static_huffman, lane QZ_TPASS);
                uint32_t hb = 0, pend = 0;
                if (btype) open_block(L.cs, L.hist, btype, gfinal, reinterpret_cast<uint32_t *>(job.slots + (size_t)g0 * job.slot_stride), lane, &hb, &pend QZ_TPASS);
                if (lane == 0) { T.btype = (uint32_t)btype; T.hb = hb; T.pend = pend; }
                QZ_MARK(11);
            }
            group_bar<QZS_TEAM * 32>(bar);
            QZ_MARK(10);            /* waiting for the leader */
            if (T.btype == 0) {
#pragma unroll
                for (int j = 0; j < QZS_PPW; j++) {
                    if (!ps[j].n) continue;
                    const uint32_t out_bytes = stored_piece(job.slots + (size_t)ps[j].g * job.slot_stride, ps[j].src, ps[j].n, ps[j].bfinal, lane);
                    if (lane == 0) job.piece_len[ps[j].g] = out_bytes;
                }
            } else {
                const uint32_t *tab = L.hist;
                uint32_t *slotw = reinterpret_cast<uint32_t *>(job.slots + (size_t)g0 * job.slot_stride);
                uint32_t beg[QZS_PPW], end[QZS_PPW], mybits[QZS_PPW], incl[QZS_PPW];
#pragma unroll
                for (int j = 0; j < QZS_PPW; j++) {
                    const uint32_t NT = T.ntok[tw * QZS_PPW + j];
                    const uint32_t R = (((NT + 31) >> 5) + 3) & ~3u;
                    beg[j] = min(lane * R, NT); end[j] = min(beg[j] + R, NT);
                    mybits[j] = count_run_bits(tab, tok_slot + (size_t)(tw * QZS_PPW + j) * QZB_TOK_STRIDE(PIECE), beg[j], end[j], pkeep);
                    incl[j] = mybits[j];
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(FULL, incl[j], o); if (lane >= (uint32_t)o) incl[j] += y; }
                    if (lane == 31) T.bits[tw * QZS_PPW + j] = incl[j];
                }
                QZ_MARK(12);
                group_bar<QZS_TEAM * 32>(bar);
                QZ_MARK(13);
                uint32_t before[QZS_PPW], total = T.hb;
#pragma unroll
                for (int j = 0; j < QZS_PPW; j++) before[j] = T.hb;
#pragma unroll
                for (int i = 0; i < QZ_GROUP; i++) {
                    const uint32_t bi = T.bits[i];
#pragma unroll
                    for (int j = 0; j < QZS_PPW; j++) if (i < (int)(tw * QZS_PPW + j)) before[j] += bi;
                    total += bi;
                }
                const uint32_t end_bit = total;
                const uint32_t nz = 3 + ((0u - (end_bit + 3)) & 7);
                const uint32_t end_bit2 = gfinal ? end_bit : end_bit + nz + 32;
#pragma unroll
                for (int j = 0; j < QZS_PPW; j++) slotw[(before[j] + incl[j] - mybits[j]) >> 5] = 0;
                if (tw == QZS_TEAM - 1 && lane == 31) slotw[end_bit2 >> 5] = 0;
                group_bar<QZS_TEAM * 32>(bar);
                QZ_MARK(14);
#pragma unroll
                for (int j = 0; j < QZS_PPW; j++) {
                    const uint32_t NT = T.ntok[tw * QZS_PPW + j];
                    const bool owns_eob = last_in_group[j] && beg[j] < NT && end[j] == NT;
                    emit_run(tab, tok_slot + (size_t)(tw * QZS_PPW + j) * QZB_TOK_STRIDE(PIECE), beg[j], end[j], before[j] + incl[j] - mybits[j], T.pend,
                             tw == 0 && j == 0 && lane == 0, !gfinal && owns_eob, nz, slotw, pkeep);
                    if (lane == 0 && ps[j].n) job.piece_len[ps[j].g] = (tw == 0 && j == 0) ? (end_bit2 + 7) >> 3 : 0u;
                }
            }
            group_bar<QZS_TEAM * 32>(bar);            /* every warp is done with the block's tokens and tables */
        }
        if (tw == 0 && lane == 0) { __threadfence_block(); S.state = QZS_FREE; }
        QZ_MARK(8);
    }
}

/* dynamic shared memory of the split kernel */
static inline size_t qzs_smem_bytes(int hb, int nmatch, int nteams)
{
    return (size_t)nmatch * (sizeof(PieceBuf<13>) + ((size_t)2 << hb)) + (size_t)nteams * QZS_TEAM * ((QZ_HIST_WORDS * 4 + 15) & ~15u) + (size_t)nteams * sizeof(GroupLead);
}
#endif
