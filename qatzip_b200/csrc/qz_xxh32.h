/* qz_xxh32.h -- xxHash32 arithmetic (reference vendored src/xxhash.c:300-437; used for the LZ4
 * frame header check byte at src/qatzip_lz4.c:130 and, on QAT, for the content checksum). */
#ifndef QZ_XXH32_H
#define QZ_XXH32_H
#include "qz_hd.h"
#define QZ_XP1 2654435761u
#define QZ_XP2 2246822519u
#define QZ_XP3 3266489917u
#define QZ_XP4 668265263u
#define QZ_XP5 374761393u
QZ_HD uint32_t qz_rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
QZ_HD uint32_t qz_xxh_round(uint32_t acc, uint32_t in) { return qz_rotl32(acc + in * QZ_XP2, 13) * QZ_XP1; }
QZ_HD uint32_t qz_xxh_avalanche(uint32_t h) { h ^= h >> 15; h *= QZ_XP2; h ^= h >> 13; h *= QZ_XP3; h ^= h >> 16; return h; }
QZ_HD uint32_t qz_xxh_rd32(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
/* one-shot scalar XXH32 (host use, and device use for the 10-byte frame descriptor) */
QZ_HD uint32_t qz_xxh32(const uint8_t *p, size_t len, uint32_t seed)
{
    const uint8_t *end = p + len; uint32_t h;
    if (len >= 16) {
        uint32_t v1 = seed + QZ_XP1 + QZ_XP2, v2 = seed + QZ_XP2, v3 = seed, v4 = seed - QZ_XP1;
        const uint8_t *lim = end - 16;
        do {
            v1 = qz_xxh_round(v1, qz_xxh_rd32(p)); v2 = qz_xxh_round(v2, qz_xxh_rd32(p + 4));
            v3 = qz_xxh_round(v3, qz_xxh_rd32(p + 8)); v4 = qz_xxh_round(v4, qz_xxh_rd32(p + 12));
            p += 16;
        } while (p <= lim);
        h = qz_rotl32(v1, 1) + qz_rotl32(v2, 7) + qz_rotl32(v3, 12) + qz_rotl32(v4, 18);
    } else h = seed + QZ_XP5;
    h += (uint32_t)len;
    while (p + 4 <= end) { h = qz_rotl32(h + qz_xxh_rd32(p) * QZ_XP3, 17) * QZ_XP4; p += 4; }
    while (p < end) { h = qz_rotl32(h + (*p++) * QZ_XP5, 11) * QZ_XP1; }
    return qz_xxh_avalanche(h);
}
#endif
