/* qz_inflate.h -- RFC 1951 decoder core as scalar host+device code.
 *
 * On the GPU one warp owns one member: lane 0 runs qz_inflate_run() (table-driven symbol
 * decode, literals stored directly), and every back-reference / stored block is handed to the
 * whole warp for the copy.  The CPU unit tests drive the same functions serially.
 *
 * Replaces the QAT stateless inflate request (reference src/qatzip.c:2191 cpaDcDecompressData
 * with CPA_DC_FLUSH_FINAL); the decoder itself lives in device firmware, not in the reference. */
#ifndef QZ_INFLATE_H
#define QZ_INFLATE_H
#include "qz_hd.h"
#include "qz_deflate_tables.h"

#define QZ_LL_LUT_BITS 9     /* 1 KiB: small enough that 64 members per SM can keep private tables in shared memory */
#define QZ_D_LUT_BITS 7

struct QzInflTables {
    uint16_t ll_lut[1 << QZ_LL_LUT_BITS];   /* (sym << 4) | len ; 0 = code longer than the LUT / unused */
    uint16_t d_lut[1 << QZ_D_LUT_BITS];
    uint16_t ll_count[16], d_count[16];      /* codes per length (slow path, canonical decode) */
    uint16_t ll_sorted[288], d_sorted[32];   /* symbols ordered by (length, symbol) */
    uint8_t lens[320];                       /* scratch: code lengths while reading a dynamic header */
};

/* LSB-first bit reader over a byte buffer of length n.  Reads past the end yield zero bits and
 * are detected afterwards through qz_br_overrun(). */
struct QzBitReader {
    const uint8_t *p;
    uint32_t n, pos;       /* pos = next byte to load */
    uint64_t acc;
    uint32_t nacc;
    uint32_t phantom;      /* zero bytes supplied beyond the end */
};
QZ_HD void qz_br_init(QzBitReader *b, const uint8_t *p, uint32_t n) { b->p = p; b->n = n; b->pos = 0; b->acc = 0; b->nacc = 0; b->phantom = 0; }
QZ_HD void qz_br_refill(QzBitReader *b)
{
    /* Top up only when half empty, and then with ONE aligned 32-bit load: the decode loop calls this
     * per symbol, so the common case must be a compare and nothing else.  Byte loads happen only
     * to reach 4-byte alignment at the start and in the last <4 bytes (then zero "phantom" bytes). */
    if (b->nacc > 32) return;
    while (b->nacc <= 56 && (((uintptr_t)(b->p + b->pos)) & 3) != 0) {
        if (b->pos < b->n) b->acc |= (uint64_t)b->p[b->pos++] << b->nacc;
        else b->phantom++;
        b->nacc += 8;
    }
    if (b->nacc > 32) return;
    if (b->pos + 4 <= b->n) { b->acc |= (uint64_t)(*(const uint32_t *)(b->p + b->pos)) << b->nacc; b->pos += 4; b->nacc += 32; return; }
    while (b->nacc <= 56) {
        if (b->pos < b->n) b->acc |= (uint64_t)b->p[b->pos++] << b->nacc;
        else b->phantom++;
        b->nacc += 8;
    }
}
QZ_HD uint32_t qz_br_bits(QzBitReader *b, uint32_t n)   /* n <= 32; caller keeps nacc >= n via refill */
{
    uint32_t v = (uint32_t)(b->acc & ((1ull << n) - 1)); b->acc >>= n; b->nacc -= n; return v;
}
/* bytes of input actually consumed (whole bytes, bits still buffered are given back) */
QZ_HD uint32_t qz_br_consumed(const QzBitReader *b) { return b->pos + b->phantom - (b->nacc >> 3); }
QZ_HD int qz_br_overrun(const QzBitReader *b) { return qz_br_consumed(b) > b->n; }

/* Build decode tables from code lengths.  Step 1 (serial): counts, canonical first codes, the
 * (length,symbol)-sorted symbol list; returns -1 for an over-subscribed set, 1 for an
 * incomplete one, 0 for a complete one.  code_of[] receives each symbol's canonical code. */
QZ_HD int qz_infl_prepare(const uint8_t *len, int n, uint16_t *count, uint16_t *sorted, uint16_t *code_of)
{
    uint16_t offs[16]; uint32_t next[16]; int left = 1;
    for (int l = 0; l < 16; l++) count[l] = 0;
    for (int s = 0; s < n; s++) count[len[s]]++;
    if (count[0] == n) return 1;
    for (int l = 1; l < 16; l++) { left <<= 1; left -= count[l]; if (left < 0) return -1; }
    offs[1] = 0; next[0] = 0; next[1] = 0;
    uint32_t code = 0;
    for (int l = 1; l < 15; l++) offs[l + 1] = (uint16_t)(offs[l] + count[l]);
    for (int l = 1; l < 16; l++) { next[l] = code; code = (code + count[l]) << 1; }
    for (int s = 0; s < n; s++) if (len[s]) { sorted[offs[len[s]]++] = (uint16_t)s; code_of[s] = (uint16_t)next[len[s]]++; }
    return left > 0 ? 1 : 0;
}
/* Step 2 (parallel over symbols: call with lane/nlanes, or 0/1 on the host): fill the LUT. */
QZ_HD void qz_infl_fill_lut(const uint8_t *len, const uint16_t *code_of, int n, uint16_t *lut, int lut_bits, int lane, int nlanes)
{
    for (int s = lane; s < n; s += nlanes) {
        uint32_t l = len[s];
        if (l == 0 || l > (uint32_t)lut_bits) continue;
        uint32_t r = qz_bitrev(code_of[s], l);
        uint16_t e = (uint16_t)((s << 4) | l);
        for (uint32_t k = r; k < (1u << lut_bits); k += (1u << l)) lut[k] = e;
    }
}
/* canonical bit-by-bit decode for codes longer than the LUT.  Returns symbol or -1. */
QZ_HD int qz_infl_slow(QzBitReader *b, const uint16_t *count, const uint16_t *sorted)
{
    int code = 0, first = 0, index = 0;
    for (int l = 1; l < 16; l++) {
        code |= (int)(b->acc & 1); b->acc >>= 1; b->nacc--;
        int c = count[l];
        if (code - c < first) return sorted[index + (code - first)];
        index += c; first += c; first <<= 1; code <<= 1;
    }
    return -1;
}

/* events returned by qz_inflate_run */
enum { QZI_MATCH = 0, QZI_END_BLOCK = 1, QZI_ERR_DATA = -1, QZI_ERR_FULL = -2, QZI_ERR_TRUNC = -3 };

/* Decode symbols of the current Huffman block: literals go straight to dst[*out], the loop
 * returns at the first back-reference (len/dist filled, NOT yet copied), at end-of-block, or
 * on error. */
QZ_HD int qz_inflate_run(QzBitReader *b, const QzInflTables *t, uint8_t *dst, uint32_t *out, uint32_t cap,
                         uint32_t *mlen, uint32_t *mdist)
{
    uint32_t o = *out;
    for (;;) {
        qz_br_refill(b);
        uint32_t e = t->ll_lut[b->acc & ((1u << QZ_LL_LUT_BITS) - 1)];
        int sym;
        if (e) { sym = (int)(e >> 4); b->acc >>= (e & 15); b->nacc -= (e & 15); }
        else { sym = qz_infl_slow(b, t->ll_count, t->ll_sorted); if (sym < 0) { *out = o; return QZI_ERR_DATA; } }
        if (sym < 256) {
            if (o >= cap) { *out = o; return QZI_ERR_FULL; }
            dst[o++] = (uint8_t)sym;
            continue;
        }
        *out = o;
        if (sym == 256) return QZI_END_BLOCK;
        sym -= 257;
        if (sym >= 29) return QZI_ERR_DATA;
        uint32_t eb, len = qz_len_base((uint32_t)sym, &eb);
        if (b->nacc < 48) qz_br_refill(b);
        len += qz_br_bits(b, eb);
        uint32_t de = t->d_lut[b->acc & ((1u << QZ_D_LUT_BITS) - 1)];
        int ds;
        if (de) { ds = (int)(de >> 4); b->acc >>= (de & 15); b->nacc -= (de & 15); }
        else { ds = qz_infl_slow(b, t->d_count, t->d_sorted); if (ds < 0) return QZI_ERR_DATA; }
        if (ds >= 30) return QZI_ERR_DATA;
        uint32_t dist = qz_dist_base((uint32_t)ds, &eb);
        dist += qz_br_bits(b, eb);
        if (dist > o) return QZI_ERR_DATA;
        if (o + len > cap) return QZI_ERR_FULL;
        *mlen = len; *mdist = dist;
        return QZI_MATCH;
    }
}

/* Batch form of the decode loop: turn the next symbols of the current Huffman block into at most
 * `max_tok` tokens (literal byte, or 1<<31 | (len-3) << 16 | (dist-1)) WITHOUT touching the
 * output; *pos is the output position before the batch and is advanced by the bytes the tokens
 * stand for.  Returns QZI_MATCH (= 0: buffer full, more to come), QZI_END_BLOCK, or an error. */
QZ_HD int qz_inflate_tokens(QzBitReader *b, const QzInflTables *t, uint32_t *tok, uint32_t max_tok, uint32_t *ntok,
                            uint32_t *pos, uint32_t cap)
{
    uint32_t n = 0, o = *pos;
    int ev = QZI_MATCH;
    while (n < max_tok) {
        qz_br_refill(b);
        uint32_t e = t->ll_lut[b->acc & ((1u << QZ_LL_LUT_BITS) - 1)];
        int sym;
        if (e) { sym = (int)(e >> 4); b->acc >>= (e & 15); b->nacc -= (e & 15); }
        else { sym = qz_infl_slow(b, t->ll_count, t->ll_sorted); if (sym < 0) { ev = QZI_ERR_DATA; break; } }
        if (sym < 256) {
            if (o >= cap) { ev = QZI_ERR_FULL; break; }
            tok[n++] = (uint32_t)sym; o++;
            continue;
        }
        if (sym == 256) { ev = QZI_END_BLOCK; break; }
        sym -= 257;
        if (sym >= 29) { ev = QZI_ERR_DATA; break; }
        uint32_t eb, len = qz_len_base((uint32_t)sym, &eb);
        if (b->nacc < 48) qz_br_refill(b);
        len += qz_br_bits(b, eb);
        uint32_t de = t->d_lut[b->acc & ((1u << QZ_D_LUT_BITS) - 1)];
        int ds;
        if (de) { ds = (int)(de >> 4); b->acc >>= (de & 15); b->nacc -= (de & 15); }
        else { ds = qz_infl_slow(b, t->d_count, t->d_sorted); if (ds < 0) { ev = QZI_ERR_DATA; break; } }
        if (ds >= 30) { ev = QZI_ERR_DATA; break; }
        uint32_t dist = qz_dist_base((uint32_t)ds, &eb);
        dist += qz_br_bits(b, eb);
        if (dist > o) { ev = QZI_ERR_DATA; break; }
        if (o + len > cap) { ev = QZI_ERR_FULL; break; }
        tok[n++] = 0x80000000u | ((len - 3) << 16) | (dist - 1);
        o += len;
    }
    /* Past the end of the input the reader supplies zero bits; a code table in which the all-zero
     * code is a length symbol would turn those into tokens for ever.  Once per batch is enough. */
    if (ev == QZI_MATCH && qz_br_overrun(b)) ev = QZI_ERR_TRUNC;
    *ntok = n; *pos = o;
    return ev;
}

/* Read the code lengths of a dynamic block header into t->lens (hlit lengths then hdist).
 * Returns 0 or QZI_ERR_DATA. */
QZ_HD int qz_inflate_read_dynamic(QzBitReader *b, QzInflTables *t, uint32_t *hlit_out, uint32_t *hdist_out)
{
    const uint8_t ORDER[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
    qz_br_refill(b);
    uint32_t hlit = qz_br_bits(b, 5) + 257, hdist = qz_br_bits(b, 5) + 1, hclen = qz_br_bits(b, 4) + 4;
    if (hlit > 286 || hdist > 30) return QZI_ERR_DATA;
    uint8_t cl[19]; uint16_t code_of[19], cl_count[16], cl_sorted[19], cl_lut[128];
    for (int i = 0; i < 19; i++) cl[i] = 0;
    for (uint32_t i = 0; i < hclen; i++) { if (b->nacc < 3) qz_br_refill(b); cl[ORDER[i]] = (uint8_t)qz_br_bits(b, 3); }
    if (qz_infl_prepare(cl, 19, cl_count, cl_sorted, code_of) != 0) return QZI_ERR_DATA;
    for (int i = 0; i < 128; i++) cl_lut[i] = 0;
    qz_infl_fill_lut(cl, code_of, 19, cl_lut, 7, 0, 1);
    uint32_t i = 0, total = hlit + hdist;
    while (i < total) {
        qz_br_refill(b);
        uint32_t e = cl_lut[b->acc & 127];
        if (!e) return QZI_ERR_DATA;
        uint32_t s = e >> 4; b->acc >>= (e & 15); b->nacc -= (e & 15);
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
