/* qz_crc32.h -- CRC-32 (IEEE 802.3, reflected) pieces shared by device and host.
 * Replaces the QAT engine's per-request checksum (reference src/qatzip_utils.c:276-281) and
 * zlib crc32_combine at reference src/qatzip.c:1707-1714. */
#ifndef QZ_CRC32_H
#define QZ_CRC32_H
#include "qz_hd.h"
#define QZ_CRC_POLY 0xEDB88320u

/* a(x)*b(x) mod P, reflected representation (bit 31 is x^0). 32 shift/xor steps. */
QZ_HD uint32_t qz_gf2_mul(uint32_t a, uint32_t b)
{
    uint32_t p = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1          /* rarely executed: keep it small in the instruction cache */
#endif
    for (int i = 0; i < 32; i++) {
        p ^= (0u - (a >> 31)) & b;
        a <<= 1;
        b = (b >> 1) ^ ((0u - (b & 1)) & QZ_CRC_POLY);
    }
    return p;
}
/* x^(8*nbytes) mod P */
QZ_HD uint32_t qz_crc_xpow8(uint64_t nbytes)
{
    uint32_t xp = 0x80000000u, sq = 0x00800000u;
    for (; nbytes; nbytes >>= 1) { if (nbytes & 1) xp = qz_gf2_mul(xp, sq); sq = qz_gf2_mul(sq, sq); }
    return xp;
}
/* crc(A||B) from crc(A), crc(B), len(B).  Both CRCs are finalised (zlib convention). */
QZ_HD uint32_t qz_crc32_combine(uint32_t crc_a, uint32_t crc_b, uint64_t len_b)
{
    return qz_gf2_mul(crc_a, qz_crc_xpow8(len_b)) ^ crc_b;
}
/* byte-table entry i (host-side table fill; device keeps the table in shared memory) */
QZ_HD uint32_t qz_crc_table_entry(uint32_t i)
{
    uint32_t c = i;
    for (int k = 0; k < 8; k++) c = (c >> 1) ^ ((0u - (c & 1)) & QZ_CRC_POLY);
    return c;
}
#endif
