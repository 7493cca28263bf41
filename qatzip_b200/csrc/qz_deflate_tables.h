/* qz_deflate_tables.h -- RFC 1951 symbol arithmetic, closed-form (no lookup tables needed on
 * the device).  Replaces the fixed-function QAT deflate engine configured at reference
 * src/qatzip_utils.c:264-341. */
#ifndef QZ_DEFLATE_TABLES_H
#define QZ_DEFLATE_TABLES_H
#include "qz_hd.h"

QZ_HD int qz_ilog2(uint32_t x)   /* floor(log2(x)), x >= 1 */
{
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)x);
#else
    return 31 - __builtin_clz(x);
#endif
}

/* match length 3..258 -> (symbol 257..285, extra-bit count, extra value) */
QZ_HD void qz_len_code(uint32_t len, uint32_t *sym, uint32_t *ebits, uint32_t *eval)
{
    uint32_t l = len - 3;
    if (l < 8) { *sym = 257 + l; *ebits = 0; *eval = 0; }
    else if (len == 258) { *sym = 285; *ebits = 0; *eval = 0; }
    else { uint32_t e = (uint32_t)qz_ilog2(l) - 2; *sym = 261 + 4 * e + ((l >> e) & 3); *ebits = e; *eval = l & ((1u << e) - 1); }
}
/* match distance 1..32768 -> (symbol 0..29, extra-bit count, extra value) */
QZ_HD void qz_dist_code(uint32_t dist, uint32_t *sym, uint32_t *ebits, uint32_t *eval)
{
    uint32_t x = dist - 1;
    if (x < 4) { *sym = x; *ebits = 0; *eval = 0; }
    else { uint32_t nb = (uint32_t)qz_ilog2(x); *sym = 2 * nb + ((x >> (nb - 1)) & 1); *ebits = nb - 1; *eval = x & ((1u << (nb - 1)) - 1); }
}
/* inverse maps used by inflate */
QZ_HD uint32_t qz_len_base(uint32_t s /*0..28*/, uint32_t *ebits)
{
    if (s < 8) { *ebits = 0; return 3 + s; }
    if (s == 28) { *ebits = 0; return 258; }
    uint32_t e = (s - 4) >> 2; *ebits = e; return 3 + ((4 + (s & 3)) << e);
}
QZ_HD uint32_t qz_dist_base(uint32_t s /*0..29*/, uint32_t *ebits)
{
    if (s < 4) { *ebits = 0; return 1 + s; }
    uint32_t e = (s >> 1) - 1; *ebits = e; return 1 + ((2 + (s & 1)) << e);
}
/* fixed-Huffman literal/length code length (RFC 1951 3.2.6) */
QZ_HD uint32_t qz_fixed_ll_len(uint32_t s) { return s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8; }

/* reverse the low `n` bits of v (deflate packs Huffman codes MSB-first into an LSB-first stream) */
QZ_HD uint32_t qz_bitrev(uint32_t v, uint32_t n)
{
#if defined(__CUDA_ARCH__)
    return __brev(v) >> (32 - n);
#else
    uint32_t r = 0; for (uint32_t i = 0; i < n; i++) { r = (r << 1) | ((v >> i) & 1); } return r;
#endif
}
#endif
