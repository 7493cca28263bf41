/* qz_adler32.h -- Adler-32 in block form (RFC 1950), host + device.
 *
 * The zlib wire format (reference src/qatzip_gzip.c:263-281: header 78 9C, footer = big-endian
 * Adler-32 of the chunk, QAT session checksum CPA_DC_ADLER32 src/qatzip_utils.c:277-283) needs a
 * checksum that a warp can compute over strips and join.  A block of n bytes is summarised by
 *     s1 = sum d[i]                 (mod 65521)
 *     s2 = sum (n - i) * d[i]       (mod 65521)
 * so that Adler-32 = ((n + s2) mod P) << 16 | ((1 + s1) mod P), and two adjacent blocks X, Y
 * join as  s1 = s1x + s1y,  s2 = s2x + len(Y) * s1x + s2y. */
#ifndef QZ_ADLER32_H
#define QZ_ADLER32_H
#include "qz_hd.h"

#define QZ_ADLER_P 65521u
#define QZ_ADLER_NMAX 5552u     /* bytes that can be summed in 32 bits between reductions */

/* sums of p[0..n): *s1, *s2 are outputs (reduced) */
QZ_HD void qz_adler_block(const uint8_t *p, uint32_t n, uint32_t *s1o, uint32_t *s2o)
{
    uint32_t s1 = 0, s2 = 0;
    while (n) {
        uint32_t k = n < QZ_ADLER_NMAX ? n : QZ_ADLER_NMAX;
        n -= k;
        while (k--) { s1 += *p++; s2 += s1; }
        s1 %= QZ_ADLER_P; s2 %= QZ_ADLER_P;
    }
    *s1o = s1; *s2o = s2;
}

/* join block X (sums s1x, s2x) with the block Y of len_y bytes that follows it */
QZ_HD void qz_adler_join(uint32_t *s1x, uint32_t *s2x, uint32_t s1y, uint32_t s2y, uint64_t len_y)
{
    const uint64_t s2 = (uint64_t)*s2x + (len_y % QZ_ADLER_P) * (uint64_t)*s1x + s2y;
    *s2x = (uint32_t)(s2 % QZ_ADLER_P);
    *s1x = (*s1x + s1y) % QZ_ADLER_P;
}

QZ_HD uint32_t qz_adler_finish(uint32_t s1, uint32_t s2, uint64_t n)
{
    const uint32_t a = (1u + s1) % QZ_ADLER_P;
    const uint32_t b = (uint32_t)((n % QZ_ADLER_P + s2) % QZ_ADLER_P);
    return b << 16 | a;
}

/* sums <-> packed word (kernels park per-piece sums in a uint32) */
QZ_HD uint32_t qz_adler_pack(uint32_t s1, uint32_t s2) { return s2 << 16 | s1; }

/* Adler-32 of X||Y from the two finished checksums (what zlib's adler32_combine computes) */
QZ_HD uint32_t qz_adler32_combine(uint32_t ax, uint32_t ay, uint64_t len_y)
{
    const uint32_t P = QZ_ADLER_P;
    const uint32_t a1 = ax & 0xffff, b1 = ax >> 16, a2 = ay & 0xffff, b2 = ay >> 16;
    const uint32_t r = (uint32_t)(len_y % P);
    const uint32_t a = (a1 + a2 + P - 1) % P;
    const uint64_t b = (uint64_t)b1 + b2 + (uint64_t)r * a1 + (uint64_t)P * P - r;   /* - r: Y's own "+len" started from A=1 */
    return (uint32_t)(b % P) << 16 | a;
}

QZ_HD uint32_t qz_adler32(const uint8_t *p, uint64_t n)
{
    uint32_t s1 = 0, s2 = 0; uint64_t done = 0;
    while (done < n) {
        const uint32_t k = (n - done) < (1u << 30) ? (uint32_t)(n - done) : (1u << 30);
        uint32_t t1, t2;
        qz_adler_block(p + done, k, &t1, &t2);
        qz_adler_join(&s1, &s2, t1, t2, k);
        done += k;
    }
    return qz_adler_finish(s1, s2, n);
}
#endif
