/* harness/corpus.c -- deterministic synthetic corpora for tests and bench.
 *
 * Two corpora, both driven by splitmix64 (never libc rand()), both addressable per 1 MiB segment
 * so any slice of a multi-GiB buffer can be produced independently and in parallel:
 *
 *   REF-RLE       restates the *distribution* of the reference's test generator
 *                 (reference test/main.c:293-310: runs of r%100 copies of byte r%65+90).
 *   SILESIA-LIKE  12-segment cycle  T X B E T X B T Z T X R  of six classes (text, markup,
 *                 binary records, executable-like, sparse, random), 1 MiB per segment.
 *
 * Build: gcc -O2 -fPIC -shared -o libqzcorpus.so corpus.c -lpthread   (see harness/Makefile)
 * This file is harness code: it is used by tests/ and bench.py, not by libqatzip.so.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <pthread.h>

#define SEG_BYTES (1u << 20)

typedef struct { uint64_t s; } rng_t;
static inline uint64_t sm64(rng_t *r)
{
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline uint32_t rnd_below(rng_t *r, uint32_t n) { return (uint32_t)((sm64(r) >> 32) * (uint64_t)n >> 32); }
static inline double rnd_unit(rng_t *r) { return (double)(sm64(r) >> 11) * (1.0 / 9007199254740992.0); }

/* ---- shared, seed-derived tables (vocabulary, zipf CDFs); built once, read-only after ---- */
#define VOCAB 4096
#define NTAGS 64
static struct {
    uint64_t seed;
    int ready;
    char word[VOCAB][13];
    uint8_t wlen[VOCAB];
    double wcdf[VOCAB];
    char tag[NTAGS][12];
    uint8_t tlen[NTAGS];
    double bcdf[256];
    uint8_t bperm[256];
} G;
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;

static void build_tables(uint64_t seed)
{
    pthread_mutex_lock(&g_lock);
    if (G.ready && G.seed == seed) { pthread_mutex_unlock(&g_lock); return; }
    rng_t r = { seed ^ 0x7461626C6573ull };
    double tot = 0;
    for (int i = 0; i < VOCAB; i++) {
        int len = 2 + (int)rnd_below(&r, 11);             /* 2..12 */
        if (i < 64) len = 1 + (int)rnd_below(&r, 4);       /* frequent words are short */
        for (int k = 0; k < len; k++) {
            /* letter frequencies skewed towards a handful of letters, like natural text */
            static const char letters[] = "eeeeeetttttaaaaooooiiiinnnnsssshhhrrrdddlllcuumwfgypbvk";
            G.word[i][k] = letters[rnd_below(&r, sizeof(letters) - 1)];
        }
        G.word[i][len] = 0;
        G.wlen[i] = (uint8_t)len;
        tot += 1.0 / pow((double)(i + 1), 1.1);
        G.wcdf[i] = tot;
    }
    for (int i = 0; i < VOCAB; i++) G.wcdf[i] /= tot;
    for (int i = 0; i < NTAGS; i++) {
        int len = 3 + (int)rnd_below(&r, 8);
        for (int k = 0; k < len; k++) G.tag[i][k] = (char)('a' + rnd_below(&r, 26));
        G.tag[i][len] = 0;
        G.tlen[i] = (uint8_t)len;
    }
    tot = 0;
    for (int i = 0; i < 256; i++) { tot += 1.0 / pow((double)(i + 1), 1.2); G.bcdf[i] = tot; G.bperm[i] = (uint8_t)i; }
    for (int i = 0; i < 256; i++) G.bcdf[i] /= tot;
    for (int i = 255; i > 0; i--) { int j = (int)rnd_below(&r, (uint32_t)i + 1); uint8_t t = G.bperm[i]; G.bperm[i] = G.bperm[j]; G.bperm[j] = t; }
    G.seed = seed;
    G.ready = 1;
    pthread_mutex_unlock(&g_lock);
}

static inline int cdf_pick(const double *cdf, int n, double u)
{
    int lo = 0, hi = n - 1;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (cdf[mid] < u) lo = mid + 1; else hi = mid; }
    return lo;
}

/* ---- class generators: each fills exactly n bytes of out ---- */
static void gen_text(rng_t *r, uint8_t *out, size_t n)
{
    size_t p = 0; int sentence = 0;
    while (p < n) {
        int w = cdf_pick(G.wcdf, VOCAB, rnd_unit(r));
        for (int k = 0; k < G.wlen[w] && p < n; k++) {
            char c = G.word[w][k];
            if (sentence == 0 && k == 0) c = (char)(c - 32);
            out[p++] = (uint8_t)c;
        }
        sentence++;
        uint32_t q = rnd_below(r, 100);
        if (p < n) {
            if (sentence > 6 && q < 12) { out[p++] = '.'; if (p < n) out[p++] = (q < 3) ? '\n' : ' '; sentence = 0; }
            else if (q < 20) { out[p++] = ','; if (p < n) out[p++] = ' '; }
            else out[p++] = ' ';
        }
    }
}

static size_t put_dec(uint8_t *out, size_t p, size_t n, uint64_t v, int mindig)
{
    char tmp[24]; int k = 0;
    do { tmp[k++] = (char)('0' + v % 10); v /= 10; } while (v || k < mindig);
    while (k && p < n) out[p++] = (uint8_t)tmp[--k];
    return p;
}
static size_t put_str(uint8_t *out, size_t p, size_t n, const char *s)
{
    while (*s && p < n) out[p++] = (uint8_t)*s++;
    return p;
}

static void gen_markup(rng_t *r, uint8_t *out, size_t n, uint64_t seg)
{
    size_t p = 0; uint64_t id = seg * 40000u + 1000u, ts = 1700000000ull + seg * 86400u;
    while (p < n) {
        int t = (int)rnd_below(r, NTAGS); if (rnd_below(r, 4)) t &= 7;   /* a few hot tags */
        p = put_str(out, p, n, "<"); p = put_str(out, p, n, G.tag[t]);
        p = put_str(out, p, n, " id=\""); p = put_dec(out, p, n, id++, 1);
        p = put_str(out, p, n, "\" ts=\""); ts += rnd_below(r, 5); p = put_dec(out, p, n, ts, 1);
        p = put_str(out, p, n, "\" v=\""); p = put_dec(out, p, n, rnd_below(r, 10), 1);
        p = put_str(out, p, n, "."); p = put_dec(out, p, n, rnd_below(r, 1000), 3);
        p = put_str(out, p, n, "\">"); p = put_str(out, p, n, G.tag[t]);
        p = put_str(out, p, n, "</"); p = put_str(out, p, n, G.tag[t]); p = put_str(out, p, n, ">\n");
    }
}

static void gen_records(rng_t *r, uint8_t *out, size_t n, uint64_t seg)
{
    uint32_t id = (uint32_t)(seg * 32768u); int32_t w[4] = { 1000, -2000, 300000, 7 };
    for (size_t p = 0; p < n; p += 32) {
        uint32_t rec[8];
        rec[0] = id++;
        for (int k = 0; k < 4; k++) { w[k] += (int32_t)rnd_below(r, 7) - 3; rec[1 + k] = (uint32_t)w[k]; }
        rec[5] = 0x3f800000u | (uint32_t)(sm64(r) & 0x7fffffu);
        rec[6] = 0; rec[7] = 0;
        size_t m = n - p < 32 ? n - p : 32;
        memcpy(out + p, rec, m);
    }
}

static void gen_exec(rng_t *r, uint8_t *out, size_t n)
{
    size_t p = 0;
    while (p < n) {
        if (p > 64 && rnd_unit(r) < 0.35 / 8.0) {          /* copy event; p=0.35 per ~8-byte unit */
            uint32_t len = 4 + rnd_below(r, 13), maxd = p < 32768 ? (uint32_t)p : 32767u;
            uint32_t d = 1 + rnd_below(r, maxd);
            for (uint32_t k = 0; k < len && p < n; k++, p++) out[p] = out[p - d];
        } else {
            out[p++] = G.bperm[cdf_pick(G.bcdf, 256, rnd_unit(r))];
        }
    }
}

static void gen_sparse(rng_t *r, uint8_t *out, size_t n)
{
    size_t p = 0;
    while (p < n) {
        uint32_t run = 64 + rnd_below(r, 8192); uint8_t v = rnd_below(r, 4) ? 0x00 : 0xFF;
        for (uint32_t k = 0; k < run && p < n; k++) out[p++] = v;
        uint32_t burst = rnd_below(r, 24);
        for (uint32_t k = 0; k < burst && p < n; k++) out[p++] = (uint8_t)sm64(r);
    }
}

static void gen_random(rng_t *r, uint8_t *out, size_t n)
{
    size_t p = 0;
    for (; p + 8 <= n; p += 8) { uint64_t v = sm64(r); memcpy(out + p, &v, 8); }
    for (; p < n; p++) out[p] = (uint8_t)sm64(r);
}

static const char CYCLE[12] = { 'T', 'X', 'B', 'E', 'T', 'X', 'B', 'T', 'Z', 'T', 'X', 'R' };

/* Fill one segment (index seg, up to 1 MiB) of the SILESIA-LIKE corpus. */
static void silesia_segment(uint64_t seed, uint64_t seg, uint8_t *out, size_t n)
{
    rng_t r = { seed * 0x9E3779B97F4A7C15ull + seg * 0xD1B54A32D192ED03ull + 1 };
    sm64(&r);
    switch (CYCLE[seg % 12]) {
    case 'T': gen_text(&r, out, n); break;
    case 'X': gen_markup(&r, out, n, seg); break;
    case 'B': gen_records(&r, out, n, seg); break;
    case 'E': gen_exec(&r, out, n); break;
    case 'Z': gen_sparse(&r, out, n); break;
    default:  gen_random(&r, out, n); break;
    }
}

static void refrle_segment(uint64_t seed, uint64_t seg, uint8_t *out, size_t n)
{
    rng_t r = { seed * 0x9E3779B97F4A7C15ull + seg * 0xD1B54A32D192ED03ull + 7 };
    size_t p = 0;
    while (p < n) {
        uint32_t j = (uint32_t)(sm64(&r) >> 33) % 100u;
        uint8_t c = (uint8_t)((uint32_t)(sm64(&r) >> 33) % 65u + 90u);
        for (uint32_t i = 0; i < j && p < n; i++) out[p++] = c;
    }
}

typedef struct { int kind; uint64_t seed; uint64_t first_seg; uint8_t *out; size_t total; int tid, nthr; } job_t;
static void *worker(void *arg)
{
    job_t *j = (job_t *)arg;
    size_t nseg = (j->total + SEG_BYTES - 1) / SEG_BYTES;
    for (size_t s = (size_t)j->tid; s < nseg; s += (size_t)j->nthr) {
        size_t off = s * SEG_BYTES, n = j->total - off < SEG_BYTES ? j->total - off : SEG_BYTES;
        if (j->kind == 0) refrle_segment(j->seed, j->first_seg + s, j->out + off, n);
        else silesia_segment(j->seed, j->first_seg + s, j->out + off, n);
    }
    return NULL;
}

/* kind 0 = REF-RLE, 1 = SILESIA-LIKE.  Fills out[0..total) with the corpus bytes that start at
 * segment index first_seg (so a 16 GiB logical buffer can be produced in 512 MiB pieces). */
int qzcorpus_fill(int kind, uint64_t seed, uint64_t first_seg, uint8_t *out, size_t total, int nthreads)
{
    if (!out || kind < 0 || kind > 1) return -1;
    build_tables(seed);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 64) nthreads = 64;
    pthread_t th[64]; job_t jobs[64];
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = (job_t){ kind, seed, first_seg, out, total, t, nthreads };
        if (t + 1 == nthreads) worker(&jobs[t]);
        else if (pthread_create(&th[t], NULL, worker, &jobs[t])) return -2;
    }
    for (int t = 0; t + 1 < nthreads; t++) pthread_join(th[t], NULL);
    return 0;
}

char qzcorpus_class_of_segment(uint64_t seg) { return CYCLE[seg % 12]; }
