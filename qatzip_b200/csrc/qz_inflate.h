/* qz_inflate.h -- RFC 1951 decoder core as scalar host+device code.
 *
 * On the GPU one warp owns one member: lane 0 turns the bit stream into tokens with
 * qz_inflate_tokens() (one table lookup per literal, two per match; bases and extra-bit counts come
 * out of the table entry), the whole warp places them.  The CPU unit tests drive the same functions
 * serially against zlib.
 *
 * Replaces the QAT stateless inflate request (reference src/qatzip.c:2191 cpaDcDecompressData
 * with CPA_DC_FLUSH_FINAL); the decoder itself lives in device firmware, not in the reference. */
#ifndef QZ_INFLATE_H
#define QZ_INFLATE_H
#include "qz_hd.h"
#include "qz_deflate_tables.h"

#ifndef QZ_LL_LUT_BITS
#define QZ_LL_LUT_BITS 10    /* 4 KiB of 32-bit entries per member being decoded */
#endif
#ifndef QZ_D_LUT_BITS
#define QZ_D_LUT_BITS 8      /* at least 7: the code-length alphabet's table borrows 128 entries */
#endif

/* Decode-table entry, laid out for the token loop's instruction count:
 *   bits 0..4   code length (1..15): a shift by the whole entry (funnel shift: low five bits) skips the code
 *   bit  5      QZE_LEN: a length or a distance symbol;  bit 6  QZE_EOB: end of block
 *   bits 8..12  code length + number of extra bits (lengths, distances): what the symbol consumes in all, a byte of its own;
 *               bits 8..15 of a literal entry hold the literal byte instead
 *   bits 16..30 length base 3..258 or distance base 1..24577: entry >> 16
 *   bit  31     QZE_LIT: a literal (the sign bit)
 * An all-zero table entry = code longer than the table, or unused.  An entry function's answer with no kind bit set = a symbol
 * that must not occur (lit/len 286, 287; distance 30, 31): such symbols are left out of the table, so the decode loop meets
 * them on its slow path only. */
#define QZE_LIT 0x80000000u
#define QZE_LEN 0x00000020u      /* also marks a valid distance entry */
#define QZE_EOB 0x00000040u
#define QZE_KIND (QZE_LIT | QZE_LEN | QZE_EOB)
#define QZE_CODE_BITS(e) ((e) & 31u)
#define QZE_ALL_BITS(e) (((e) >> 8) & 0xffu)
#define QZE_BASE(e) ((e) >> 16)

/* tokens per batch of the decode loop, and the compressed words staged for one batch (see the token loop below) */
#ifndef QZ_INFL_BATCH
#define QZ_INFL_BATCH 64                    /* tokens per batch: a multiple of 32 (the warp places 32 at a time) */
#endif
#define QZ_INFL_INW (2 * QZ_INFL_BATCH)     /* staged words: bits 0 .. 31 + BATCH * 48 and the word after, rounded up to whole lanes */

struct QzInflTables {
    uint32_t ll_lut[1 << QZ_LL_LUT_BITS];
    uint32_t d_lut[1 << QZ_D_LUT_BITS];
    uint16_t ll_count[16], d_count[16];      /* codes per length */
    uint16_t ll_first[16], d_first[16];      /* canonical first code of each length */
    uint16_t ll_offs[16], d_offs[16];        /* index in sorted[] of the first symbol of each length */
    uint16_t ll_sorted[288], d_sorted[32];   /* symbols ordered by (length, symbol) */
    union {
        uint8_t lens[320];                   /* while a block's tables are built: code lengths read from its header (hlit, then hdist) */
        uint32_t stage[QZ_INFL_INW];         /* while the block is decoded: the compressed words the next batch can reach */
    };
};

QZ_HD uint32_t qz_infl_ll_entry(uint32_t s, uint32_t l)
{
    if (s < 256) return QZE_LIT | (s << 8) | l;
    if (s == 256) return QZE_EOB | l;
    if (s > 285) return l;
    uint32_t eb, base = qz_len_base(s - 257, &eb);
    return QZE_LEN | (base << 16) | ((l + eb) << 8) | l;
}
QZ_HD uint32_t qz_infl_d_entry(uint32_t s, uint32_t l)
{
    if (s > 29) return l;
    uint32_t eb, base = qz_dist_base(s, &eb);
    return QZE_LEN | (base << 16) | ((l + eb) << 8) | l;
}

/* LSB-first bit reader.  Input is fetched as aligned 32-bit words, one word ahead of need (wnext),
 * so the decode loop never waits for the load it has just issued.  Offsets are relative to the aligned
 * address at or below the first byte; bytes past the end read as zero and are detected afterwards
 * through qz_br_overrun(). */
struct QzBitReader {
    const uint8_t *base;   /* p rounded down to 4 bytes */
    uint32_t skew;         /* p - base */
    uint32_t end;          /* skew + n */
    uint32_t n;
    uint32_t pos;          /* offset of wnext */
    uint32_t wnext;
    uint64_t acc;
    uint32_t nacc;
};
QZ_HD uint32_t qz_br_word(const QzBitReader *b, uint32_t off)
{
    if (off + 4 <= b->end) return *(const uint32_t *)(b->base + off);
    uint32_t w = 0;
    for (uint32_t k = 0; k < 4; k++) if (off + k < b->end) w |= (uint32_t)b->base[off + k] << (8 * k);
    return w;
}
/* position the reader at byte `off` of the input */
QZ_HD void qz_br_seek(QzBitReader *b, uint32_t off)
{
    const uint32_t a = off + b->skew, w = a & ~3u, sub = a & 3u;
    b->acc = (uint64_t)(qz_br_word(b, w) >> (8 * sub));
    b->nacc = 32 - 8 * sub;
    b->pos = w + 4;
    b->wnext = qz_br_word(b, b->pos);
}
QZ_HD void qz_br_init(QzBitReader *b, const uint8_t *p, uint32_t n)
{
    b->skew = (uint32_t)((uintptr_t)p & 3); b->base = p - b->skew; b->end = b->skew + n; b->n = n;
    qz_br_seek(b, 0);
}
/* after this at least 33 bits are buffered */
QZ_HD void qz_br_refill(QzBitReader *b)
{
    if (b->nacc > 32) return;
    b->acc |= (uint64_t)b->wnext << b->nacc;
    b->nacc += 32;
    b->pos += 4;
    b->wnext = qz_br_word(b, b->pos);
}
QZ_HD uint32_t qz_br_bits(QzBitReader *b, uint32_t n)   /* n <= 32; caller keeps nacc >= n via refill */
{
    uint32_t v = (uint32_t)(b->acc & ((1ull << n) - 1)); b->acc >>= n; b->nacc -= n; return v;
}
/* bytes of input actually consumed (whole bytes; bits still buffered are given back) */
QZ_HD uint32_t qz_br_consumed(const QzBitReader *b) { return b->pos - b->skew - (b->nacc >> 3); }
QZ_HD int qz_br_overrun(const QzBitReader *b) { return qz_br_consumed(b) > b->n; }
/* nothing but zero fill is left (QZ_DEFLATE_RAW chunks that end without BFINAL stop here) */
QZ_HD int qz_br_exhausted(const QzBitReader *b) { return (uint64_t)(b->pos - b->skew) * 8 >= (uint64_t)b->n * 8 + b->nacc; }

/* Build the canonical-code description of one alphabet from its code lengths (serial form; the
 * kernel has a warp-parallel equivalent): counts, first codes, offsets, and the (length, symbol)-
 * sorted symbol list.  Returns -1 for an over-subscribed set, 1 for an incomplete one, 0 otherwise. */
QZ_HD int qz_infl_prepare(const uint8_t *len, int n, uint16_t *count, uint16_t *first, uint16_t *offs, uint16_t *sorted)
{
    uint16_t fill[16]; int left = 1;
    for (int l = 0; l < 16; l++) { count[l] = 0; first[l] = 0; offs[l] = 0; }
    for (int s = 0; s < n; s++) count[len[s]]++;
    if (count[0] == n) { count[0] = 0; return 1; }
    for (int l = 1; l < 16; l++) { left <<= 1; left -= count[l]; if (left < 0) return -1; }
    uint32_t code = 0; offs[0] = 0; first[0] = 0; offs[1] = 0;
    for (int l = 1; l < 16; l++) { first[l] = (uint16_t)code; code = (code + count[l]) << 1; if (l < 15) offs[l + 1] = (uint16_t)(offs[l] + count[l]); }
    for (int l = 0; l < 16; l++) fill[l] = offs[l];
    for (int s = 0; s < n; s++) if (len[s]) sorted[fill[len[s]]++] = (uint16_t)s;
    return left > 0 ? 1 : 0;
}
/* Fill the lookup table (parallel over sorted symbols: call with lane/nlanes, or 0/1 on the host).
 * The table must be zero beforehand. */
QZ_HD void qz_infl_fill_lut(const uint8_t *len, const uint16_t *count, const uint16_t *first, const uint16_t *offs, const uint16_t *sorted,
                            uint32_t *lut, int lut_bits, int is_dist, int lane, int nlanes)
{
    const int used = offs[15] + count[15];
    for (int i = lane; i < used; i += nlanes) {
        const uint32_t s = sorted[i], l = len[s];
        if (l > (uint32_t)lut_bits) continue;
        const uint32_t r = qz_bitrev((uint32_t)first[l] + (uint32_t)(i - offs[l]), l);
        const uint32_t e = is_dist ? qz_infl_d_entry(s, l) : qz_infl_ll_entry(s, l);
        if (!(e & QZE_KIND)) continue;
        for (uint32_t k = r; k < (1u << lut_bits); k += (1u << l)) lut[k] = e;
    }
}
/* Canonical decode for codes longer than the table's `lut_bits` (the table has already ruled out every shorter code): the
 * next 15 bits as an MSB-first number; its l-bit prefix is a code of length l exactly when it falls into that length's range
 * of consecutive codes.  Returns the symbol and its code length in *len_out, or -1. */
QZ_HD int qz_infl_slow(uint64_t acc, const uint16_t *count, const uint16_t *first, const uint16_t *offs, const uint16_t *sorted, int lut_bits, uint32_t *len_out)
{
    const uint32_t rev = qz_bitrev((uint32_t)acc & 0x7fffu, 15);
    for (int l = lut_bits + 1; l < 16; l++) {
        const uint32_t idx = (rev >> (15 - l)) - first[l];
        if (idx < count[l]) { *len_out = (uint32_t)l; return sorted[offs[l] + idx]; }
    }
    return -1;
}

/* events returned by qz_inflate_tokens */
enum { QZI_MATCH = 0, QZI_END_BLOCK = 1, QZI_ERR_DATA = -1, QZI_ERR_FULL = -2, QZI_ERR_TRUNC = -3 };

/* Token produced by the decode loop: a literal is its table entry as is (QZE_LIT set, byte in bits 8..15),
 * a match is len << 16 | dist (len 3..258, dist 1..32768) with bit 31 clear. */
QZ_HD int qz_tok_is_literal(uint32_t t) { return (t >> 31) != 0; }
QZ_HD uint32_t qz_tok_byte(uint32_t t) { return (t >> 8) & 0xff; }
QZ_HD uint32_t qz_tok_len(uint32_t t) { return (t >> 16) & 0x1ff; }
QZ_HD uint32_t qz_tok_dist(uint32_t t) { return t & 0xffff; }

/* ---- the token loop ----
 * Turns the next symbols of the current Huffman block into at most QZ_INFL_BATCH tokens WITHOUT touching the output.  This is
 * the serial heart of inflate (one lane per member runs it) and it is bound by instruction issue, so it is written for
 * instruction count: the compressed words a batch can reach (32 tokens of at most 48 bits, starting less than 32 bits into
 * the first word) are STAGED in a small window beforehand (by the whole warp on the device), and inside the loop the reader is
 * nothing but a bit offset into that window -- a symbol's bits are two loads and a funnel shift away, consuming them is one
 * add, and there is no refill code at all.  The output position is not stepped per literal (every token stands for at least
 * one byte: position = adj + token count), and away from the end of the destination the room checks are compiled out. */
#define QZ_INFL_ROOMY (QZ_INFL_BATCH * 258u) /* with this much room left no batch can overflow the destination */

/* where the reader stands: byte offset (from b->base) of the word that holds the next bit, and the bit's offset in that word */
QZ_HD void qz_br_where(const QzBitReader *b, uint32_t *woff, uint32_t *sh)
{
    const uint32_t held = (b->nacc + 31) >> 5;                               /* words the accumulator reaches back over */
    *sh = held * 32 - b->nacc; *woff = b->pos - 4 * held;
}
/* the stream's word at byte offset off from base (bytes past `end` read as zero) */
QZ_HD uint32_t qz_word_at(const uint8_t *base, uint32_t end, uint32_t off)
{
    if (off + 4 <= end) return *(const uint32_t *)(base + off);
    uint32_t w = 0;
    for (uint32_t k = 0; k < 4; k++) if (off + k < end) w |= (uint32_t)base[off + k] << (8 * k);
    return w;
}
/* put the reader at bit lp of the window that was staged from byte offset woff */
QZ_HD void qz_br_resume(QzBitReader *b, uint32_t woff, const uint32_t *inw, uint32_t lp)
{
    const uint32_t wi = lp >> 5, s = lp & 31;
    b->acc = (uint64_t)(inw[wi] >> s); b->nacc = 32 - s; b->pos = woff + 4 * wi + 4; b->wnext = inw[wi + 1];
}
/* the 32 bits that start `sh` (low five bits used) bits into lo, continuing in hi: one SHF on the device */
QZ_HD uint32_t qz_funnel(uint32_t lo, uint32_t hi, uint32_t sh)
{
#ifdef __CUDA_ARCH__
    return __funnelshift_r(lo, hi, sh);
#else
    return (uint32_t)(((((uint64_t)hi) << 32) | lo) >> (sh & 31));
#endif
}

/* the value of a length or distance symbol: base + the extra bits that follow its code in `bits` */
QZ_HD uint32_t qz_sym_value(uint32_t e, uint32_t bits, uint32_t all)
{
    const uint32_t field = bits & ~(0xffffffffu << all);         /* code and extra bits (all <= 28) */
#ifdef __CUDA_ARCH__
    return QZE_BASE(e) + __funnelshift_r(field, 0u, e);          /* the shift takes the entry's low five bits: the code length */
#else
    return QZE_BASE(e) + (field >> QZE_CODE_BITS(e));
#endif
}

/* CAREFUL = true: the destination may fill up in this batch -- every token is checked against the room left and against the
 * output position (a distance must not reach before the start), *pos is advanced by the bytes the tokens stand for.
 * CAREFUL = false needs cap - *pos >= QZ_INFL_ROOMY: no batch can overflow, so the loop does not follow the output position at
 * all; *pos is left alone, and the distances are checked against the positions when the tokens are placed (the placing side
 * computes every token's position anyway).  *lp: bit offset into inw, in and out.  Returns QZI_MATCH (= 0: batch full, more
 * to come), QZI_END_BLOCK, or an error. */
template <bool CAREFUL>
QZ_HD int qz_inflate_tokens_core(const uint32_t *inw, uint32_t *lp_io, const QzInflTables *t, uint32_t *tok, uint32_t *ntok, uint32_t *pos, uint32_t cap)
{
    int ev = QZI_MATCH;
    uint32_t lp = *lp_io, nt = 0, adj = *pos, nmax = QZ_INFL_BATCH;          /* CAREFUL: output position = adj + nt */
    const uint32_t *const ll_lut = t->ll_lut, *const d_lut = t->d_lut;
    if (CAREFUL) {
        const uint32_t room = cap - adj;
        if (room < nmax) nmax = room;
        if (room == 0) {
            /* the output is full: the block may still end here, anything else does not fit */
            const uint32_t bits = qz_funnel(inw[lp >> 5], inw[(lp >> 5) + 1], lp);
            uint32_t e = ll_lut[bits & ((1u << QZ_LL_LUT_BITS) - 1)];
            if (e == 0) {
                uint32_t l; const int sym = qz_infl_slow(bits, t->ll_count, t->ll_first, t->ll_offs, t->ll_sorted, QZ_LL_LUT_BITS, &l);
                e = sym < 0 ? 0u : qz_infl_ll_entry((uint32_t)sym, l);
            }
            if ((int32_t)e >= 0 && (e & QZE_EOB)) { lp += QZE_CODE_BITS(e); ev = QZI_END_BLOCK; }
            else ev = ((int32_t)e < 0 || (e & QZE_LEN)) ? QZI_ERR_FULL : QZI_ERR_DATA;
            goto done;
        }
    }
    while (nt != nmax) {
        uint32_t wi = lp >> 5;
        uint32_t bits = qz_funnel(inw[wi], inw[wi + 1], lp);
        uint32_t e = ll_lut[bits & ((1u << QZ_LL_LUT_BITS) - 1)];
        if ((int32_t)e < 0) {                                /* literal */
lit:
            lp += QZE_CODE_BITS(e);
            tok[nt++] = e;
            continue;
        }
        if (!(e & QZE_LEN)) {
            /* off the fast path: a code longer than the table, the end of the block, or a symbol that must not occur */
            if (e == 0) {
                uint32_t l; const int sym = qz_infl_slow(bits, t->ll_count, t->ll_first, t->ll_offs, t->ll_sorted, QZ_LL_LUT_BITS, &l);
                if (sym < 0) { ev = QZI_ERR_DATA; goto done; }
                e = qz_infl_ll_entry((uint32_t)sym, l);
                if ((int32_t)e < 0) goto lit;
            }
            if (!(e & QZE_LEN)) {
                if (e & QZE_EOB) { lp += QZE_CODE_BITS(e); ev = QZI_END_BLOCK; }
                else ev = QZI_ERR_DATA;
                goto done;
            }
        }
        {   /* length: code + extra bits, at most 20 of the 32 in hand; distance: at most 28 */
            const uint32_t all = QZE_ALL_BITS(e);
            const uint32_t len = qz_sym_value(e, bits, all);
            lp += all;
            wi = lp >> 5;
            bits = qz_funnel(inw[wi], inw[wi + 1], lp);
            uint32_t de = d_lut[bits & ((1u << QZ_D_LUT_BITS) - 1)];
            if (!(de & QZE_LEN)) {
                uint32_t l = 0; const int ds = qz_infl_slow(bits, t->d_count, t->d_first, t->d_offs, t->d_sorted, QZ_D_LUT_BITS, &l);
                if (ds < 0) { ev = QZI_ERR_DATA; goto done; }
                de = qz_infl_d_entry((uint32_t)ds, l);
                if (!(de & QZE_LEN)) { ev = QZI_ERR_DATA; goto done; }
            }
            const uint32_t dall = QZE_ALL_BITS(de);
            const uint32_t dist = qz_sym_value(de, bits, dall);
            lp += dall;
            if (CAREFUL) {
                const uint32_t o = adj + nt;
                if (dist > o) { ev = QZI_ERR_DATA; goto done; }
                if (len > cap - o) { ev = QZI_ERR_FULL; goto done; }      /* (o <= cap always; no 32-bit wrap) */
                adj += len - 1;
                const uint32_t left = cap - o - len;                 /* tokens still to come each need a byte of it */
                if (nmax - (nt + 1) > left) nmax = nt + 1 + left;
            }
            tok[nt++] = (len << 16) | dist;
        }
    }
done:
    *lp_io = lp; *ntok = nt;
    if (CAREFUL) *pos = adj + nt;
    return ev;
}

/* the whole step on one thread (host tests; the kernel stages with all lanes, calls the core itself and learns the new
 * output position from placing the tokens) */
QZ_HD int qz_inflate_tokens(QzBitReader *b, const QzInflTables *t, uint32_t *tok, uint32_t *ntok, uint32_t *pos, uint32_t cap)
{
    uint32_t inw[QZ_INFL_INW], woff, lp;
    qz_br_where(b, &woff, &lp);
    for (uint32_t i = 0; i < QZ_INFL_INW; i++) inw[i] = qz_word_at(b->base, b->end, woff + 4 * i);
    int ev;
    if (cap - *pos < QZ_INFL_ROOMY) ev = qz_inflate_tokens_core<true>(inw, &lp, t, tok, ntok, pos, cap);
    else {
        ev = qz_inflate_tokens_core<false>(inw, &lp, t, tok, ntok, pos, cap);
        uint32_t o = *pos;
        for (uint32_t k = 0; k < *ntok; k++) {
            if (qz_tok_is_literal(tok[k])) o++;
            else { if (qz_tok_dist(tok[k]) > o) { *ntok = k; ev = QZI_ERR_DATA; break; } o += qz_tok_len(tok[k]); }
        }
        *pos = o;
    }
    qz_br_resume(b, woff, inw, lp);
    /* Past the end of the input the reader supplies zero bits; a code table in which the all-zero
     * code is a length symbol would turn those into tokens for ever.  Once per batch is enough. */
    if (ev == QZI_MATCH && qz_br_overrun(b)) ev = QZI_ERR_TRUNC;
    return ev;
}

/* Read the code lengths of a dynamic block header into t->lens (hlit lengths then hdist).  The
 * code-length alphabet's own tables borrow the (not yet built) distance table.  Returns 0 or QZI_ERR_DATA. */
QZ_HD int qz_inflate_read_dynamic(QzBitReader *b, QzInflTables *t, uint32_t *hlit_out, uint32_t *hdist_out)
{
    const uint8_t ORDER[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
    qz_br_refill(b);
    uint32_t hlit = qz_br_bits(b, 5) + 257, hdist = qz_br_bits(b, 5) + 1, hclen = qz_br_bits(b, 4) + 4;
    if (hlit > 286 || hdist > 30) return QZI_ERR_DATA;
    uint8_t cl[19]; uint16_t cl_count[16], cl_first[16], cl_offs[16], cl_sorted[19];
    uint32_t *cl_lut = t->d_lut;                             /* 128 of its 256 entries */
    for (int i = 0; i < 19; i++) cl[i] = 0;
    for (uint32_t i = 0; i < hclen; i++) { qz_br_refill(b); cl[ORDER[i]] = (uint8_t)qz_br_bits(b, 3); }
    if (qz_infl_prepare(cl, 19, cl_count, cl_first, cl_offs, cl_sorted) != 0) return QZI_ERR_DATA;
    for (int i = 0; i < 128; i++) cl_lut[i] = 0;
    /* entries: symbol << 8 | length (QZE_LIT layout without the flag) */
    { const int used = cl_offs[15] + cl_count[15];
      for (int i = 0; i < used; i++) { const uint32_t s = cl_sorted[i], l = cl[s];
          const uint32_t r = qz_bitrev((uint32_t)cl_first[l] + (uint32_t)(i - cl_offs[l]), l);
          for (uint32_t k = r; k < 128; k += (1u << l)) cl_lut[k] = (s << 8) | l; } }
    uint32_t i = 0, total = hlit + hdist;
    while (i < total) {
        qz_br_refill(b);
        const uint32_t e = cl_lut[b->acc & 127];
        if (!e) return QZI_ERR_DATA;
        const uint32_t s = e >> 8; b->acc >>= (e & 15); b->nacc -= (e & 15);
        if (s < 16) { t->lens[i++] = (uint8_t)s; continue; }
        uint32_t rep; uint8_t v = 0;
        if (s == 16) { if (i == 0) return QZI_ERR_DATA; v = t->lens[i - 1]; rep = 3 + qz_br_bits(b, 2); }
        else if (s == 17) rep = 3 + qz_br_bits(b, 3);
        else rep = 11 + qz_br_bits(b, 7);
        if (i + rep > total) return QZI_ERR_DATA;
        while (rep--) t->lens[i++] = v;
    }
    if (t->lens[256] == 0) return QZI_ERR_DATA;
    *hlit_out = hlit; *hdist_out = hdist;
    return 0;
}
/* code lengths of the fixed block type (RFC 1951 3.2.6): 288 literal/length + 30 distance */
QZ_HD void qz_inflate_fixed_lens(QzInflTables *t)
{
    for (int s = 0; s < 288; s++) t->lens[s] = (uint8_t)qz_fixed_ll_len((uint32_t)s);
    for (int s = 0; s < 30; s++) t->lens[288 + s] = 5;
}
#endif
