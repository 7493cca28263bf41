/* qz_huffman.h -- Huffman code construction and deflate dynamic-block header coding, written as
 * scalar host+device code: on the GPU one lane of the piece's warp runs it on shared-memory
 * arrays; on the CPU the unit tests run the very same source against zlib's inflate.
 *
 * Replaces: the dynamic-Huffman tree generation inside the QAT deflate engine (reference
 * src/qatzip_utils.c:270-275 only *selects* CPA_DC_HT_FULL_DYNAMIC / CPA_DC_HT_STATIC and
 * auto-select-best; the construction itself is device firmware and is not in the reference). */
#ifndef QZ_HUFFMAN_H
#define QZ_HUFFMAN_H
#include "qz_hd.h"
#include "qz_deflate_tables.h"

#define QZ_NUM_LL 286
#define QZ_NUM_D 30
#define QZ_NUM_CL 19
#define QZ_HUFF_KEY(freq, sym) (((uint32_t)(freq) << 9) | (uint32_t)(sym))

/* LSB-first bit writer over 32-bit words (slot buffers are 4-byte aligned). */
struct QzBitWriter {
    uint32_t *words;
    uint64_t acc;
    uint32_t nacc;     /* valid bits in acc, always < 32 between calls */
    uint32_t wpos;     /* next word index */
};
QZ_HD void qz_bw_init(QzBitWriter *bw, uint32_t *words) { bw->words = words; bw->acc = 0; bw->nacc = 0; bw->wpos = 0; }
QZ_HD void qz_bw_put(QzBitWriter *bw, uint32_t bits, uint32_t n)   /* n <= 32 */
{
    bw->acc |= (uint64_t)bits << bw->nacc;
    bw->nacc += n;
    if (bw->nacc >= 32) { bw->words[bw->wpos++] = (uint32_t)bw->acc; bw->acc >>= 32; bw->nacc -= 32; }
}
QZ_HD uint32_t qz_bw_bitpos(const QzBitWriter *bw) { return bw->wpos * 32 + bw->nacc; }
/* zero-pad to the next byte boundary */
QZ_HD void qz_bw_align_byte(QzBitWriter *bw) { uint32_t r = bw->nacc & 7; if (r) qz_bw_put(bw, 0, 8 - r); }
/* write out the (<32) pending bits as a final partial word; returns total length in bytes */
QZ_HD uint32_t qz_bw_finish(QzBitWriter *bw)
{
    uint32_t bytes = bw->wpos * 4 + ((bw->nacc + 7) >> 3);
    if (bw->nacc) bw->words[bw->wpos] = (uint32_t)bw->acc;
    return bytes;
}

/* deflate requires every alphabet to carry at least two codes (zlib's inflate rejects a
 * one-code code-length alphabet): give unused low symbols a count of 1 until two are in use. */
QZ_HD void qz_huff_force_two(uint32_t *freq, int n)
{
    QZ_ASSUME_SHARED(freq);
    int used = 0;
    for (int i = 0; i < n && used < 2; i++) used += freq[i] != 0;
    for (int i = 0; used < 2 && i < n; i++) if (freq[i] == 0) { freq[i] = 1; used++; }
}

/* In-place minimum-redundancy code lengths (Moffat & Katajainen, 1995) over frequencies sorted
 * ascending in A[0..n), n >= 2, in three passes.  Passes 1-2 (tree by two-queue merge, then depths of
 * the n-1 internal nodes, root = node n-2 at depth 0, depths non-increasing with the index) are serial;
 * pass 3 (internal depths -> leaf depths) has a serial form here and a warp-parallel one in the kernel.
 * On return of the full routine A[i] is the code length of the i-th sorted symbol (non-increasing in i). */
/* pass 1 alone: on return A[i], i < n - 2, is the parent of internal node i; the root is node n - 2 */
QZ_HD_SERIAL void qz_huff_merge_pass(uint32_t *A, int n)
{
    /* The two-queue merge.  This loop is the longest single-lane stretch of the compressor (seven warps of a
     * group wait for it), so it is written for its dependency chain: the weights at the heads of both queues and one
     * element behind each head live in registers, every load is issued one consumption ahead of its use, and the
     * choices are selects rather than branches (the literal/length and the distance tree run in two lanes of one
     * warp and stay converged).  Same tree as the textbook form: ties go to the leaf.
     *   ar  = weight of internal node `root`      (INF while the queue is empty, i.e. root == next)
     *   ar2 = weight of internal node `root + 1`  (INF while that node does not exist yet)
     *   al, al2 = weights of leaves `leaf`, `leaf + 1` (INF past the end)
     * In-place safety: leaf + root = 2 next throughout, so writes to A[next] stay below every unread leaf. */
    QZ_ASSUME_SHARED(A);
    const uint32_t INF = 0xffffffffu;
    int root = 0, leaf = 2, next;
    uint32_t ar = A[0] + A[1], ar2 = INF;
    uint32_t al = leaf < n ? A[leaf] : INF, al2 = leaf + 1 < n ? A[leaf + 1] : INF;
    A[0] = ar;
    for (next = 1; next < n - 1; next++) {
        uint32_t s = 0;
#pragma unroll
        for (int child = 0; child < 2; child++) {
            const bool ti = ar < al;                        /* internal node is strictly lighter; an empty queue never is */
            s += ti ? ar : al;
            if (ti) A[root] = (uint32_t)next;               /* parent pointer */
            root += ti ? 1 : 0; leaf += ti ? 0 : 1;
            const int la = ti ? root + 1 : leaf + 1;        /* the element that becomes the new look-ahead */
            const uint32_t ld = (ti ? la < next : la < n) ? A[la] : INF;
            if (ti) { ar = ar2; ar2 = ld; } else { al = al2; al2 = ld; }
        }
        A[next] = s;
        /* the node just made may be the head of the internal queue, or right behind it */
        if (root == next) ar = s; else if (root + 1 == next) ar2 = s;
    }
}
QZ_HD_SERIAL void qz_huff_inplace_depths(uint32_t *A, int n)
{
    QZ_ASSUME_SHARED(A);
    qz_huff_merge_pass(A, n);
    A[n - 2] = 0;
    for (int next = n - 3; next >= 0; next--) A[next] = A[A[next]] + 1;
}
QZ_HD_SERIAL void qz_huff_depths_to_lengths(uint32_t *A, int n)
{
    QZ_ASSUME_SHARED(A);
    int avbl = 1, used = 0, dpth = 0, root = n - 2, next = n - 1;
    while (avbl > 0) {
        while (root >= 0 && (int)A[root] == dpth) { used++; root--; }
        while (avbl > used) { A[next--] = (uint32_t)dpth; avbl--; }
        avbl = 2 * used; dpth++; used = 0;
    }
}
QZ_HD_SERIAL void qz_huff_inplace_lengths(uint32_t *A, int n)
{
    qz_huff_inplace_depths(A, n);
    qz_huff_depths_to_lengths(A, n);
}

/* Cap the (sorted, non-increasing) code lengths in len[0..n) at maxbits, keeping the Kraft sum
 * exact: fold over-long codes to maxbits, then repeatedly turn the deepest leaf above the cap
 * into an internal node that adopts one folded leaf; finally hand lengths back out longest
 * first (= least frequent first). */
QZ_HD_SERIAL void qz_huff_limit_sorted(uint32_t *len, int n, int maxbits)
{
    QZ_ASSUME_SHARED(len);
    if ((int)len[0] <= maxbits) return;
    uint32_t cnt[16];
    for (int l = 0; l < 16; l++) cnt[l] = 0;
    for (int i = 0; i < n; i++) cnt[(int)len[i] > maxbits ? maxbits : (int)len[i]]++;
    uint32_t kraft = 0;
    for (int l = 1; l <= maxbits; l++) kraft += cnt[l] << (maxbits - l);
    while (kraft > (1u << maxbits)) {
        int b = maxbits - 1;
        while (cnt[b] == 0) b--;
        cnt[b]--; cnt[b + 1] += 2; cnt[maxbits]--;
        kraft--;
    }
    int i = 0;
    for (int l = maxbits; l >= 1; l--) for (uint32_t c = cnt[l]; c; c--) len[i++] = (uint32_t)l;
}

/* Sorted keys (QZ_HUFF_KEY(freq, sym), ascending, all freq > 0, n_used >= 2) -> per-symbol code
 * lengths capped at maxbits.  keys[] is clobbered, ids[] is scratch for n_used entries.
 * len_by_sym[] must be zero for symbols that do not occur.  (The kernel runs the first and the
 * last loop across the warp and only the two calls in between on one lane.) */
QZ_HD_SERIAL void qz_huff_lengths_from_sorted(uint32_t *keys, uint16_t *ids, int n_used, int maxbits, uint8_t *len_by_sym)
{
    QZ_ASSUME_SHARED(keys); QZ_ASSUME_SHARED(ids); QZ_ASSUME_SHARED(len_by_sym);
    for (int i = 0; i < n_used; i++) { ids[i] = (uint16_t)(keys[i] & 511u); keys[i] >>= 9; }
    qz_huff_inplace_lengths(keys, n_used);
    qz_huff_limit_sorted(keys, n_used, maxbits);
    for (int i = 0; i < n_used; i++) len_by_sym[ids[i]] = (uint8_t)keys[i];
}

/* Canonical codes, already bit-reversed for LSB-first emission: out[s] = code | len << 16. */
QZ_HD_SERIAL void qz_huff_codes(const uint8_t *len, int n, uint32_t *out)
{
    uint32_t cnt[16], next[17];
    for (int l = 0; l < 16; l++) cnt[l] = 0;
    for (int s = 0; s < n; s++) cnt[len[s]]++;
    cnt[0] = 0; next[0] = 0; next[1] = 0;
    for (int l = 1; l < 16; l++) next[l + 1] = (next[l] + cnt[l]) << 1;
    for (int s = 0; s < n; s++) {
        uint32_t l = len[s];
        out[s] = l ? (qz_bitrev(next[l]++, l) | (l << 16)) : 0u;
    }
}

/* ---- dynamic block header (RFC 1951 3.2.7) ---- */
struct QzDynHeaderCore {
    uint16_t items[QZ_NUM_LL + QZ_NUM_D];   /* sym | extra_value << 5 | extra_bits_count << 12 */
    uint32_t nitems;
    uint32_t hlit, hdist, hclen;
    uint8_t cl_len[QZ_NUM_CL];
    uint32_t bits;                          /* total header bits including the 3-bit block header */
};
/* the serial planner keeps its run-length scratch inside the header; the kernel's warp-parallel planner
 * borrows (dead) sort-key space instead and only carries the core */
struct QzDynHeader : QzDynHeaderCore {
    uint8_t seq[QZ_NUM_LL + QZ_NUM_D + 4];  /* scratch: the code lengths being run-length coded */
};

QZ_HD uint16_t qz_cl_item(uint32_t sym, uint32_t eval, uint32_t ebits) { return (uint16_t)(sym | (eval << 5) | (ebits << 12)); }

/* Code-length alphabet (<= 19 symbols, 7-bit cap) from its frequencies cf[]; sizes the header.
 * h->items / h->nitems / h->hlit / h->hdist must already be set.  scratch: 32 words in the same memory as cf and h
 * (shared memory in the kernels -- every pointer the device-side callers of these routines pass is). */
QZ_HD_SERIAL void qz_cl_build(uint32_t *cf, QzDynHeaderCore *h, uint32_t *scratch)
{
    QZ_ASSUME_SHARED(cf); QZ_ASSUME_SHARED(h); QZ_ASSUME_SHARED(scratch);
    /* every item costs its symbol's code plus that symbol's extra bits, so the header size follows from
     * the counters alone (taken before two-code forcing adds symbols that are never written) */
    uint32_t unused = 0;
    for (int k = 0; k < QZ_NUM_CL; k++) unused |= (uint32_t)(cf[k] == 0) << k;
    qz_huff_force_two(cf, QZ_NUM_CL);
    uint32_t *keys = scratch; uint16_t *ids = (uint16_t *)(scratch + 20); int nu = 0;
    for (int k = 0; k < QZ_NUM_CL; k++) {
        h->cl_len[k] = 0;
        if (cf[k]) {
            uint32_t key = QZ_HUFF_KEY(cf[k], k); int p = nu++;
            while (p > 0 && keys[p - 1] > key) { keys[p] = keys[p - 1]; p--; }
            keys[p] = key;
        }
    }
    qz_huff_lengths_from_sorted(keys, ids, nu, 7, h->cl_len);
    const uint8_t ORDER[QZ_NUM_CL] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
    uint32_t hclen = QZ_NUM_CL;
    while (hclen > 4 && h->cl_len[ORDER[hclen - 1]] == 0) hclen--;
    h->hclen = hclen;
    uint32_t bits = 3 + 5 + 5 + 4 + 3 * hclen;
    for (int k = 0; k < QZ_NUM_CL; k++) if (!((unused >> k) & 1)) bits += cf[k] * (h->cl_len[k] + (k == 16 ? 2u : k == 17 ? 3u : k == 18 ? 7u : 0u));
    h->bits = bits;
}

/* Run-length code the two length arrays into code-length symbols and size the header. */
QZ_HD_SERIAL void qz_dyn_header_plan(const uint8_t *ll_len, const uint8_t *d_len, QzDynHeader *h)
{
    uint32_t hlit = QZ_NUM_LL, hdist = QZ_NUM_D;
    while (hlit > 257 && ll_len[hlit - 1] == 0) hlit--;
    while (hdist > 1 && d_len[hdist - 1] == 0) hdist--;
    h->hlit = hlit; h->hdist = hdist;
    /* the two length arrays back to back, so the run-length scan below walks one plain array */
    uint8_t *seq = h->seq;
    for (uint32_t k = 0; k < hlit; k++) seq[k] = ll_len[k];
    for (uint32_t k = 0; k < hdist; k++) seq[hlit + k] = d_len[k];
    uint32_t total = hlit + hdist, i = 0, n = 0, cf[QZ_NUM_CL];
    for (int k = 0; k < QZ_NUM_CL; k++) cf[k] = 0;
    while (i < total) {
        const uint32_t v = seq[i];
        uint32_t j = i + 1;
        while (j < total && seq[j] == v) j++;
        uint32_t run = j - i;
        if (v == 0) {
            while (run >= 11) { uint32_t r = run > 138 ? 138 : run; h->items[n++] = qz_cl_item(18, r - 11, 7); cf[18]++; run -= r; }
            if (run >= 3) { h->items[n++] = qz_cl_item(17, run - 3, 3); cf[17]++; run = 0; }
            while (run) { h->items[n++] = qz_cl_item(0, 0, 0); cf[0]++; run--; }
        } else {
            h->items[n++] = qz_cl_item(v, 0, 0); cf[v]++; run--;
            while (run >= 3) { uint32_t r = run > 6 ? 6 : run; h->items[n++] = qz_cl_item(16, r - 3, 2); cf[16]++; run -= r; }
            while (run) { h->items[n++] = qz_cl_item(v, 0, 0); cf[v]++; run--; }
        }
        i = j;
    }
    h->nitems = n;
    uint32_t scratch[32];
    qz_cl_build(cf, h, scratch);
}

/* the fixed part of a dynamic block header: BFINAL/BTYPE, HLIT, HDIST, HCLEN and the 3-bit lengths */
QZ_HD_SERIAL void qz_dyn_header_write_prefix(QzBitWriter *bw, const QzDynHeaderCore *h, int bfinal)
{
    const uint8_t ORDER[QZ_NUM_CL] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
    qz_bw_put(bw, (uint32_t)(bfinal ? 1 : 0) | (2u << 1), 3);
    qz_bw_put(bw, h->hlit - 257, 5);
    qz_bw_put(bw, h->hdist - 1, 5);
    qz_bw_put(bw, h->hclen - 4, 4);
    for (uint32_t k = 0; k < h->hclen; k++) qz_bw_put(bw, h->cl_len[ORDER[k]], 3);
}

QZ_HD_SERIAL void qz_dyn_header_write(QzBitWriter *bw, const QzDynHeader *h, int bfinal)
{
    const uint8_t ORDER[QZ_NUM_CL] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
    uint32_t cl_code[QZ_NUM_CL];
    qz_huff_codes(h->cl_len, QZ_NUM_CL, cl_code);
    qz_bw_put(bw, (uint32_t)(bfinal ? 1 : 0) | (2u << 1), 3);
    qz_bw_put(bw, h->hlit - 257, 5);
    qz_bw_put(bw, h->hdist - 1, 5);
    qz_bw_put(bw, h->hclen - 4, 4);
    for (uint32_t k = 0; k < h->hclen; k++) qz_bw_put(bw, h->cl_len[ORDER[k]], 3);
    for (uint32_t k = 0; k < h->nitems; k++) {
        uint32_t it = h->items[k], c = cl_code[it & 31];
        qz_bw_put(bw, c & 0xffff, c >> 16);
        if (it >> 12) qz_bw_put(bw, (it >> 5) & 127, it >> 12);
    }
}
#endif
