/* oracle/qz_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the chunked compress/decompress hot path of intel/QATzip, used
 * exclusively as the checker in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg.  Nothing under oracle/ is linked into, loaded by, or called from libqatzip.so.
 *
 * What is restated here (reference file:line cited at each function):
 *   - per-chunk framing exactly as the reference's hardware path stitches it
 *     (src/qatzip.c:1610-1764 doCompressOut; src/qatzip_gzip.c:98-143,228-237;
 *      src/qatzip_lz4.c:104-143)
 *   - member/frame discovery + verification on decode
 *     (src/qatzip_utils.c:1232-1345 checkHeader; :1483-1532 decompOutCheckSum)
 *   - CRC-32 (zlib crc32 / crc32_combine semantics used at src/qatzip.c:1707-1714)
 *   - xxHash32 (vendored src/xxhash.c:300-437, only XXH32 is used: src/qatzip_lz4.c:130)
 *   - RFC 1951 inflate and the LZ4 block format (published formats; decode side restated in
 *     full so decoded bytes never depend on the library being tested).
 *
 * Third-party arithmetic that is NOT under /root/reference and is NOT restated: the DEFLATE
 * *compressor*.  The reference's device is a QAT ASIC whose exact bits are unspecified and its
 * CPU path calls zlib (src/qatzip_sw.c:147,197; zlib "≥1.2.7", this image: 1.3).  The port
 * therefore calls zlib's deflate() for chunk payloads.  Compressed bytes are unpinned by
 * design; decoded bytes, checksums, sizes and framing are pinned.
 *
 * Pinning: tests/test_oracle.py checks this file against oracle/_ref (the unmodified reference
 * sources compiled by oracle/Makefile) in both directions and against tests/golden/ fixtures
 * generated from oracle/_ref, plus zlib's crc32()/inflate() and gzip(1).
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#include <zlib.h>

#define QZ_OK 0
#define QZ_PARAMS (-1)
#define QZ_FAIL (-2)
#define QZ_BUF_ERROR (-3)
#define QZ_DATA_ERROR (-4)

enum { FMT_4B = 0, FMT_GZIP = 1, FMT_GZIP_EXT = 2, FMT_RAW = 3, FMT_LZ4 = 4, FMT_ZLIB = 5 };

static inline uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
static inline void wr32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }

/* ------------------------------------------------------------------ CRC-32 (IEEE, reflected) */
static uint32_t g_crc_tab[256];
static int g_crc_ready;
static void crc_init(void)
{
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        g_crc_tab[i] = c;
    }
    g_crc_ready = 1;
}
/* Same contract as zlib crc32(crc, buf, len): pass 0 to start. */
uint32_t qzo_crc32(uint32_t crc, const uint8_t *p, size_t n)
{
    if (!g_crc_ready) crc_init();
    crc = ~crc;
    while (n--) crc = g_crc_tab[(crc ^ *p++) & 0xff] ^ (crc >> 8);
    return ~crc;
}
/* a(x)*b(x) mod P in the reflected representation (bit 31 = x^0). */
static uint32_t gf2_mul(uint32_t a, uint32_t b)
{
    uint32_t p = 0;
    for (uint32_t m = 0x80000000u; m; m >>= 1) {
        if (a & m) p ^= b;
        b = (b & 1) ? (b >> 1) ^ 0xEDB88320u : b >> 1;
    }
    return p;
}
/* crc(A||B) from crc(A), crc(B), len(B) -- the operation at reference src/qatzip.c:1711. */
uint32_t qzo_crc32_combine(uint32_t crc_a, uint32_t crc_b, uint64_t len_b)
{
    uint32_t xp = 0x80000000u;            /* x^0 */
    uint32_t sq = 0x00800000u;            /* x^8 : one byte */
    for (uint64_t n = len_b; n; n >>= 1) { if (n & 1) xp = gf2_mul(xp, sq); sq = gf2_mul(sq, sq); }
    return gf2_mul(crc_a, xp) ^ crc_b;
}

/* ------------------------------------------------------------------ xxHash32 (src/xxhash.c) */
#define P1 2654435761u
#define P2 2246822519u
#define P3 3266489917u
#define P4 668265263u
#define P5 374761393u
static inline uint32_t rotl(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
/* reference src/xxhash.c:306-312 (round), :404-437 (stripe loop), :328-400 (finalize), :315-323 (avalanche) */
uint32_t qzo_xxh32(const uint8_t *p, size_t len, uint32_t seed)
{
    const uint8_t *end = p + len; uint32_t h;
    if (len >= 16) {
        uint32_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
        const uint8_t *lim = end - 16;
        do {
            v1 = rotl(v1 + rd32(p) * P2, 13) * P1; v2 = rotl(v2 + rd32(p + 4) * P2, 13) * P1;
            v3 = rotl(v3 + rd32(p + 8) * P2, 13) * P1; v4 = rotl(v4 + rd32(p + 12) * P2, 13) * P1;
            p += 16;
        } while (p <= lim);
        h = rotl(v1, 1) + rotl(v2, 7) + rotl(v3, 12) + rotl(v4, 18);
    } else h = seed + P5;
    h += (uint32_t)len;
    while (p + 4 <= end) { h = rotl(h + rd32(p) * P3, 17) * P4; p += 4; }
    while (p < end) { h = rotl(h + (*p++) * P5, 11) * P1; }
    h ^= h >> 15; h *= P2; h ^= h >> 13; h *= P3; h ^= h >> 16;
    return h;
}

/* ------------------------------------------------------------------ RFC 1951 inflate (restated) */
typedef struct { const uint8_t *in; size_t inlen, inpos; uint64_t bits; int nbits; } bitrd_t;
static inline void br_fill(bitrd_t *b) { while (b->nbits <= 56 && b->inpos < b->inlen) { b->bits |= (uint64_t)b->in[b->inpos++] << b->nbits; b->nbits += 8; } }
static inline int br_get(bitrd_t *b, int n, uint32_t *v)
{
    if (b->nbits < n) { br_fill(b); if (b->nbits < n) return -1; }
    *v = (uint32_t)(b->bits & ((1ull << n) - 1)); b->bits >>= n; b->nbits -= n; return 0;
}
typedef struct { uint16_t count[16]; uint16_t sym[288]; } huff_t;
static int huff_build(huff_t *h, const uint8_t *len, int n)
{
    uint16_t offs[16]; int left = 1;
    memset(h->count, 0, sizeof(h->count));
    for (int i = 0; i < n; i++) h->count[len[i]]++;
    if (h->count[0] == n) return 0;                      /* no codes: legal, decode will fail if used */
    for (int l = 1; l < 16; l++) { left <<= 1; left -= h->count[l]; if (left < 0) return -1; }
    offs[1] = 0;
    for (int l = 1; l < 15; l++) offs[l + 1] = (uint16_t)(offs[l] + h->count[l]);
    for (int i = 0; i < n; i++) if (len[i]) h->sym[offs[len[i]]++] = (uint16_t)i;
    return left;                                          /* >0: incomplete set */
}
static int huff_decode(bitrd_t *b, const huff_t *h)
{
    int code = 0, first = 0, index = 0;
    for (int l = 1; l < 16; l++) {
        uint32_t bit; if (br_get(b, 1, &bit)) return -1;
        code |= (int)bit;
        int cnt = h->count[l];
        if (code - cnt < first) return h->sym[index + (code - first)];
        index += cnt; first += cnt; first <<= 1; code <<= 1;
    }
    return -2;
}
static const uint16_t LBASE[29] = { 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258 };
static const uint8_t LEXT[29] = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0 };
static const uint16_t DBASE[30] = { 1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577 };
static const uint8_t DEXT[30] = { 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13 };

/* Inflate one raw deflate stream.  Stops after the block with BFINAL=1, or -- when
 * stop_at_input_end != 0 -- cleanly at a block boundary once the input is exhausted (the
 * QZ_DEFLATE_RAW non-final chunks end with a byte-aligned empty stored block and no BFINAL:
 * reference src/qatzip_utils.c:1082-1087).
 * Returns 0 ok, -1 corrupt, -2 output full, -3 input truncated. */
int qzo_inflate_raw(const uint8_t *src, size_t n, uint8_t *dst, size_t cap, size_t *consumed,
                    size_t *produced, int stop_at_input_end, int *saw_final)
{
    bitrd_t b = { src, n, 0, 0, 0 }; size_t out = 0; uint32_t v; int final = 0;
    static const uint8_t ORDER[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
    if (saw_final) *saw_final = 0;
    while (!final) {
        if (stop_at_input_end) {
            br_fill(&b);
            if (b.nbits < 8 && b.inpos >= b.inlen && b.nbits == 0) break;   /* clean end between blocks */
        }
        if (br_get(&b, 1, &v)) return -3;
        final = (int)v;
        if (br_get(&b, 2, &v)) return -3;
        if (v == 0) {
            uint32_t len, nlen; int drop = b.nbits & 7;
            b.bits >>= drop; b.nbits -= drop;
            if (br_get(&b, 16, &len) || br_get(&b, 16, &nlen)) return -3;
            if ((len ^ 0xffffu) != nlen) return -1;
            for (uint32_t i = 0; i < len; i++) { if (br_get(&b, 8, &v)) return -3; if (out >= cap) return -2; dst[out++] = (uint8_t)v; }
        } else if (v == 1 || v == 2) {
            huff_t hl, hd; uint8_t lens[320];
            if (v == 1) {
                int i = 0; for (; i < 144; i++) lens[i] = 8; for (; i < 256; i++) lens[i] = 9;
                for (; i < 280; i++) lens[i] = 7; for (; i < 288; i++) lens[i] = 8;
                huff_build(&hl, lens, 288);
                for (i = 0; i < 30; i++) lens[i] = 5;
                huff_build(&hd, lens, 30);
            } else {
                uint32_t nl, nd, nc; huff_t hc; uint8_t cl[19] = { 0 };
                if (br_get(&b, 5, &nl) || br_get(&b, 5, &nd) || br_get(&b, 4, &nc)) return -3;
                nl += 257; nd += 1; nc += 4;
                if (nl > 286 || nd > 30) return -1;
                for (uint32_t i = 0; i < nc; i++) { if (br_get(&b, 3, &v)) return -3; cl[ORDER[i]] = (uint8_t)v; }
                if (huff_build(&hc, cl, 19) != 0) return -1;
                uint32_t i = 0;
                while (i < nl + nd) {
                    int s = huff_decode(&b, &hc);
                    if (s < 0) return s == -1 ? -3 : -1;
                    if (s < 16) lens[i++] = (uint8_t)s;
                    else {
                        uint32_t rep; uint8_t val = 0;
                        if (s == 16) { if (i == 0) return -1; val = lens[i - 1]; if (br_get(&b, 2, &rep)) return -3; rep += 3; }
                        else if (s == 17) { if (br_get(&b, 3, &rep)) return -3; rep += 3; }
                        else { if (br_get(&b, 7, &rep)) return -3; rep += 11; }
                        if (i + rep > nl + nd) return -1;
                        while (rep--) lens[i++] = val;
                    }
                }
                if (lens[256] == 0) return -1;
                int r = huff_build(&hl, lens, (int)nl);
                if (r < 0 || (r > 0 && nl - hl.count[0] != 1)) return -1;
                r = huff_build(&hd, lens + nl, (int)nd);
                if (r < 0 || (r > 0 && nd - hd.count[0] != 1)) return -1;
            }
            for (;;) {
                int s = huff_decode(&b, &hl);
                if (s < 0) return s == -1 ? -3 : -1;
                if (s < 256) { if (out >= cap) return -2; dst[out++] = (uint8_t)s; }
                else if (s == 256) break;
                else {
                    s -= 257; if (s >= 29) return -1;
                    uint32_t eb; if (br_get(&b, LEXT[s], &eb)) return -3;
                    uint32_t len = LBASE[s] + eb;
                    int ds = huff_decode(&b, &hd);
                    if (ds < 0) return ds == -1 ? -3 : -1;
                    if (ds >= 30) return -1;
                    if (br_get(&b, DEXT[ds], &eb)) return -3;
                    uint32_t dist = DBASE[ds] + eb;
                    if (dist > out) return -1;
                    if (out + len > cap) return -2;
                    for (uint32_t k = 0; k < len; k++, out++) dst[out] = dst[out - dist];
                }
            }
        } else return -1;
    }
    /* give back whole unread bytes */
    size_t back = (size_t)(b.nbits >> 3);
    if (consumed) *consumed = b.inpos - back;
    if (produced) *produced = out;
    if (saw_final) *saw_final = final;
    return 0;
}

/* ------------------------------------------------------------------ LZ4 block format (restated) */
/* Greedy single-probe encoder obeying the block end rules (last 5 bytes literal, last match
 * starts >= 12 bytes before the end, offsets <= 65535).  Returns bytes written, 0 if it did not fit. */
size_t qzo_lz4_block_compress(const uint8_t *src, size_t n, uint8_t *dst, size_t cap)
{
    enum { HB = 13 };
    static __thread uint32_t tab[1 << HB];
    size_t ip = 0, anchor = 0, op = 0;
    memset(tab, 0xff, sizeof(tab));
    size_t mflimit = n >= 12 ? n - 12 : 0, matchlimit = n >= 5 ? n - 5 : 0;
    while (n >= 13 && ip < mflimit) {
        uint32_t seq = rd32(src + ip), h = (seq * 2654435761u) >> (32 - HB), cand = tab[h];
        tab[h] = (uint32_t)ip;
        if (cand != 0xffffffffu && ip - cand <= 65535 && rd32(src + cand) == seq) {
            size_t ml = 4; while (ip + ml < matchlimit && src[cand + ml] == src[ip + ml]) ml++;
            size_t ll = ip - anchor, need = 1 + ll / 255 + 1 + ll + 2 + (ml - 4) / 255 + 1;
            if (op + need > cap) return 0;
            uint8_t *tok = dst + op++; *tok = (uint8_t)((ll >= 15 ? 15 : ll) << 4);
            if (ll >= 15) { size_t r = ll - 15; for (; r >= 255; r -= 255) dst[op++] = 255; dst[op++] = (uint8_t)r; }
            memcpy(dst + op, src + anchor, ll); op += ll;
            dst[op++] = (uint8_t)(ip - cand); dst[op++] = (uint8_t)((ip - cand) >> 8);
            size_t mc = ml - 4; *tok |= (uint8_t)(mc >= 15 ? 15 : mc);
            if (mc >= 15) { size_t r = mc - 15; for (; r >= 255; r -= 255) dst[op++] = 255; dst[op++] = (uint8_t)r; }
            ip += ml; anchor = ip;
        } else ip++;
    }
    size_t ll = n - anchor;
    if (op + 1 + ll / 255 + 1 + ll > cap) return 0;
    uint8_t *tok = dst + op++; *tok = (uint8_t)((ll >= 15 ? 15 : ll) << 4);
    if (ll >= 15) { size_t r = ll - 15; for (; r >= 255; r -= 255) dst[op++] = 255; dst[op++] = (uint8_t)r; }
    memcpy(dst + op, src + anchor, ll); op += ll;
    return op;
}
/* Decode one LZ4 block; `hist` bytes of earlier output precede dst (linked blocks). Returns
 * produced bytes or -1 on malformed input / overflow. */
long qzo_lz4_block_decompress(const uint8_t *src, size_t n, uint8_t *dst, size_t cap, size_t hist)
{
    size_t ip = 0, op = 0;
    while (ip < n) {
        uint8_t tok = src[ip++]; size_t ll = tok >> 4;
        if (ll == 15) { uint8_t s; do { if (ip >= n) return -1; s = src[ip++]; ll += s; } while (s == 255); }
        if (ip + ll > n || op + ll > cap) return -1;
        memcpy(dst + op, src + ip, ll); ip += ll; op += ll;
        if (ip >= n) break;
        if (ip + 2 > n) return -1;
        size_t off = src[ip] | (size_t)src[ip + 1] << 8; ip += 2;
        if (off == 0 || off > op + hist) return -1;
        size_t ml = tok & 15;
        if (ml == 15) { uint8_t s; do { if (ip >= n) return -1; s = src[ip++]; ml += s; } while (s == 255); }
        ml += 4;
        if (op + ml > cap) return -1;
        for (size_t k = 0; k < ml; k++, op++) dst[op] = dst[(long)op - (long)off];
    }
    return (long)op;
}

/* ------------------------------------------------------------------ framing (hardware-path layout) */
static size_t hdr_sz(int fmt) { return fmt == FMT_GZIP_EXT ? 24 : fmt == FMT_GZIP ? 10 : fmt == FMT_4B ? 4 : fmt == FMT_LZ4 ? 15 : fmt == FMT_ZLIB ? 2 : 0; }
static size_t ftr_sz(int fmt) { return (fmt == FMT_GZIP_EXT || fmt == FMT_GZIP || fmt == FMT_LZ4) ? 8 : fmt == FMT_ZLIB ? 4 : 0; }

/* Adler-32 (RFC 1950), the checksum of zlib-format sessions: reference src/qatzip_utils.c:277-283 */
uint32_t qzo_adler32(uint32_t adler, const uint8_t *p, size_t n)
{
    uint32_t a = adler & 0xffff, b = adler >> 16;
    while (n) {
        size_t k = n < 5552 ? n : 5552; n -= k;
        while (k--) { a += *p++; b += a; }
        a %= 65521u; b %= 65521u;
    }
    return b << 16 | a;
}

/* reference src/qatzip_gzip.c:98-143 (gzip-ext / gzip / 4B), src/qatzip_lz4.c:104-132 (LZ4 frame) */
static void gen_header(int fmt, uint8_t *p, uint32_t consumed, uint32_t produced)
{
    static const uint8_t std[10] = { 0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 0xff };
    switch (fmt) {
    case FMT_GZIP: memcpy(p, std, 10); break;
    case FMT_GZIP_EXT:
        memcpy(p, std, 10); p[3] = 4; p[10] = 12; p[11] = 0; p[12] = 'Q'; p[13] = 'Z'; p[14] = 8; p[15] = 0;
        wr32(p + 16, consumed); wr32(p + 20, produced); break;
    case FMT_4B: wr32(p, produced); break;
    case FMT_ZLIB: p[0] = 0x78; p[1] = 0x9C; break;                /* reference src/qatzip_gzip.c:263-271 */
    case FMT_LZ4:
        wr32(p, 0x184D2204u); p[4] = 0x4C; p[5] = 0x40; wr32(p + 6, consumed); wr32(p + 10, 0);
        p[14] = (uint8_t)(qzo_xxh32(p + 4, 10, 0) >> 8); break;
    default: break;
    }
}
/* reference src/qatzip_gzip.c:228-237, src/qatzip_lz4.c:134-143 */
static void gen_footer(int fmt, uint8_t *p, uint32_t checksum, uint32_t consumed)
{
    if (fmt == FMT_GZIP || fmt == FMT_GZIP_EXT) { wr32(p, checksum); wr32(p + 4, consumed); }
    else if (fmt == FMT_LZ4) { wr32(p, 0); wr32(p + 4, checksum); }
    else if (fmt == FMT_ZLIB) { p[0] = (uint8_t)(checksum >> 24); p[1] = (uint8_t)(checksum >> 16); p[2] = (uint8_t)(checksum >> 8); p[3] = (uint8_t)checksum; }   /* htonl: src/qatzip_gzip.c:273-281 */
}

/* DEST_SZ: reference src/qatzip_internal.h:99 */
static size_t dest_sz(size_t n) { return (9 * n + 7) / 8 + 1024; }

/* One chunk -> raw deflate with zlib (third party), ended FINAL or FULL-flushed as the QAT
 * flush flag would (reference src/qatzip_utils.c:1082-1087). */
static int deflate_chunk(const uint8_t *src, size_t n, uint8_t *dst, size_t cap, int level, int final, size_t *produced)
{
    z_stream z; memset(&z, 0, sizeof(z));
    if (deflateInit2(&z, level, Z_DEFLATED, -15, 9, Z_DEFAULT_STRATEGY) != Z_OK) return QZ_FAIL;
    z.next_in = (Bytef *)src; z.avail_in = (uInt)n; z.next_out = dst; z.avail_out = (uInt)cap;
    int r = deflate(&z, final ? Z_FINISH : Z_FULL_FLUSH);
    int ok = final ? (r == Z_STREAM_END) : (r == Z_OK && z.avail_in == 0 && z.avail_out > 0);
    *produced = z.total_out;
    deflateEnd(&z);
    return ok ? QZ_OK : QZ_BUF_ERROR;
}

/* Chunked compress, one self-contained member/frame per hw_buff_sz chunk
 * (reference src/qatzip.c:1483-1604 doCompressIn + :1610-1764 doCompressOut).
 * *crc accumulates like reference :1707-1714 (0 restarts).  Partial progress on a full dest
 * returns QZ_BUF_ERROR with whole chunks only (src/qatzip_utils.c:1192). */
int qzo_compress(int fmt, int level, uint32_t hw_buff_sz, const uint8_t *src, size_t *src_len, uint8_t *dst,
                 size_t *dst_len, int last, uint32_t *crc)
{
    size_t in = 0, out = 0, n = *src_len, cap = *dst_len; int rc = QZ_OK;
    if (fmt < 0 || fmt > FMT_ZLIB || hw_buff_sz < 1024 || (hw_buff_sz & (hw_buff_sz - 1))) return QZ_PARAMS;
    uint8_t *tmp = (uint8_t *)malloc(dest_sz(hw_buff_sz) + 64);
    if (!tmp) return QZ_FAIL;
    do {
        size_t send = n - in < hw_buff_sz ? n - in : hw_buff_sz, produced = 0; uint32_t cks;
        int is_last_chunk = (in + send == n);
        if (fmt == FMT_LZ4) {
            size_t c = send ? qzo_lz4_block_compress(src + in, send, tmp + 4, send - 1 < dest_sz(send) ? (send ? send - 1 : 0) : dest_sz(send)) : 0;
            if (send == 0) produced = 0;
            else if (c == 0) { wr32(tmp, (uint32_t)send | 0x80000000u); memcpy(tmp + 4, src + in, send); produced = send + 4; }
            else { wr32(tmp, (uint32_t)c); produced = c + 4; }
            cks = qzo_xxh32(src + in, send, 0);
        } else {
            int final = (fmt != FMT_RAW) || (is_last_chunk && last);
            rc = deflate_chunk(src + in, send, tmp, dest_sz(hw_buff_sz), level, final, &produced);
            if (rc != QZ_OK) break;
            cks = (fmt == FMT_ZLIB) ? qzo_adler32(1, src + in, send) : qzo_crc32(0, src + in, send);
        }
        if (out + hdr_sz(fmt) + produced + ftr_sz(fmt) > cap) { rc = QZ_BUF_ERROR; break; }
        gen_header(fmt, dst + out, (uint32_t)send, (uint32_t)produced); out += hdr_sz(fmt);
        memcpy(dst + out, tmp, produced); out += produced;
        gen_footer(fmt, dst + out, cks, (uint32_t)send); out += ftr_sz(fmt);
        if (crc && fmt != FMT_LZ4 && fmt != FMT_ZLIB) *crc = (*crc == 0) ? cks : qzo_crc32_combine(*crc, cks, send);
        in += send;
    } while (in < n);
    free(tmp);
    *src_len = in; *dst_len = out;
    return rc;
}

/* Walk members/frames and decode each (reference src/qatzip.c:2103-2404 + checkHeader
 * src/qatzip_utils.c:1232-1345 + decompOutCheckSum :1483-1532).  Unlike the reference's
 * hardware path there is no hw_buff_sz cap here: oversize members are what the reference hands
 * to its software path, which yields the same bytes. */
int qzo_decompress(int fmt, const uint8_t *src, size_t *src_len, uint8_t *dst, size_t *dst_len)
{
    size_t in = 0, out = 0, n = *src_len, cap = *dst_len; int rc = QZ_OK;
    while (in < n) {
        const uint8_t *p = src + in; size_t avail = n - in, consumed = 0, produced = 0; int fin = 0, r;
        if (avail < hdr_sz(fmt)) { rc = QZ_DATA_ERROR; break; }
        if (fmt == FMT_GZIP || fmt == FMT_GZIP_EXT) {
            if (p[0] != 0x1f || p[1] != 0x8b || p[2] != 8) { rc = QZ_FAIL; break; }
            size_t h = 10; uint8_t flg = p[3];
            if (flg & 4) { if (avail < 12) { rc = QZ_DATA_ERROR; break; } h += 2 + (p[10] | (size_t)p[11] << 8); }
            if (flg & 8) { while (h < avail && p[h]) h++; h++; }
            if (flg & 16) { while (h < avail && p[h]) h++; h++; }
            if (flg & 2) h += 2;
            if (h >= avail) { rc = QZ_DATA_ERROR; break; }
            r = qzo_inflate_raw(p + h, avail - h, dst + out, cap - out, &consumed, &produced, 0, &fin);
            if (r == -2) { rc = QZ_BUF_ERROR; break; }
            if (r == -3) { rc = QZ_DATA_ERROR; break; }
            if (r) { rc = QZ_DATA_ERROR; break; }
            if (h + consumed + 8 > avail) { rc = QZ_DATA_ERROR; break; }
            if (rd32(p + h + consumed) != qzo_crc32(0, dst + out, produced) || rd32(p + h + consumed + 4) != (uint32_t)produced) { rc = QZ_DATA_ERROR; break; }
            in += h + consumed + 8; out += produced;
        } else if (fmt == FMT_ZLIB) {
            /* header test: reference src/qatzip_gzip.c:283-306; trailer: big-endian Adler-32 */
            if ((p[0] & 0x0f) != 8 || (p[0] >> 4) > 7 || (p[1] & 0x20) || ((unsigned)p[0] * 256 + p[1]) % 31) { rc = QZ_FAIL; break; }
            r = qzo_inflate_raw(p + 2, avail - 2, dst + out, cap - out, &consumed, &produced, 0, &fin);
            if (r == -2) { rc = QZ_BUF_ERROR; break; }
            if (r) { rc = QZ_DATA_ERROR; break; }
            if (2 + consumed + 4 > avail) { rc = QZ_DATA_ERROR; break; }
            const uint8_t *f = p + 2 + consumed;
            if (((uint32_t)f[0] << 24 | (uint32_t)f[1] << 16 | (uint32_t)f[2] << 8 | f[3]) != qzo_adler32(1, dst + out, produced)) { rc = QZ_DATA_ERROR; break; }
            in += 2 + consumed + 4; out += produced;
        } else if (fmt == FMT_4B) {
            uint32_t blk = rd32(p);
            if (4 + (size_t)blk > avail) { rc = QZ_DATA_ERROR; break; }
            r = qzo_inflate_raw(p + 4, blk, dst + out, cap - out, &consumed, &produced, 0, &fin);
            if (r == -2) { rc = QZ_BUF_ERROR; break; }
            if (r) { rc = QZ_DATA_ERROR; break; }
            in += 4 + blk; out += produced;
        } else if (fmt == FMT_RAW) {
            r = qzo_inflate_raw(p, avail, dst + out, cap - out, &consumed, &produced, 1, &fin);
            if (r == -2) { rc = QZ_BUF_ERROR; break; }
            if (r) { rc = QZ_DATA_ERROR; break; }
            in += consumed; out += produced;
            if (consumed == 0) break;
        } else {
            if (rd32(p) != 0x184D2204u) { rc = QZ_FAIL; break; }
            uint8_t flg = p[4]; size_t h = 6, start = out;
            if ((flg >> 6) != 1) { rc = QZ_FAIL; break; }
            if (flg & 8) h += 8;
            if (flg & 1) h += 4;
            h += 1;
            if (h > avail) { rc = QZ_DATA_ERROR; break; }
            if (p[h - 1] != (uint8_t)(qzo_xxh32(p + 4, h - 5, 0) >> 8)) { rc = QZ_DATA_ERROR; break; }
            int bad = 0;
            for (;;) {
                if (h + 4 > avail) { bad = QZ_DATA_ERROR; break; }
                uint32_t bh = rd32(p + h); h += 4;
                if (bh == 0) break;
                uint32_t bs = bh & 0x7fffffffu;
                if (h + bs > avail) { bad = QZ_DATA_ERROR; break; }
                if (bh & 0x80000000u) { if (out + bs > cap) { bad = QZ_BUF_ERROR; break; } memcpy(dst + out, p + h, bs); out += bs; }
                else {
                    long d = qzo_lz4_block_decompress(p + h, bs, dst + out, cap - out, (flg & 0x20) ? 0 : out - start);
                    if (d < 0) { bad = (cap - out < 65536) ? QZ_BUF_ERROR : QZ_DATA_ERROR; break; }
                    out += (size_t)d;
                }
                h += bs;
                if (flg & 0x10) h += 4;
            }
            if (bad) { out = start; rc = bad; break; }
            if (flg & 4) {
                if (h + 4 > avail) { out = start; rc = QZ_DATA_ERROR; break; }
                if (rd32(p + h) != qzo_xxh32(dst + start, out - start, 0)) { out = start; rc = QZ_DATA_ERROR; break; }
                h += 4;
            }
            if ((flg & 8) && rd32(p + 6) != (uint32_t)(out - start)) { out = start; rc = QZ_DATA_ERROR; break; }
            in += h;
        }
    }
    *src_len = in; *dst_len = out;
    return rc;
}
