// tools/lzmodel.cpp -- design-space model (NOT product, NOT oracle): estimates the compressed size a
// GPU-shaped LZ77+Huffman scheme would reach, to pick parameters before writing the kernel.
// Build: g++ -O2 -o /tmp/lzmodel tools/lzmodel.cpp -lz
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <queue>
#include <string>
#include <zlib.h>
#include <dlfcn.h>

struct Params {
    int piece = 65536;     // window reset granularity
    int sub = 16384;       // deflate block granularity (own huffman tables)
    int W = 1024;          // positions looked up before any of them is inserted
    int hb = 14;           // hash bits
    int hbytes = 4;        // bytes hashed
    int minm = 4;          // min match
    int capA = 258;        // cap on extension in parallel phase (model: final len anyway)
    int S = 1024;          // super-tile: matches truncated at S boundaries (0 = off)
    int rle = 1;           // also try distance 1
    int warp = 1;          // intra-warp same-hash candidate (match_any)
    int winner = 1;        // 1: highest p in window wins table slot, 0: lowest
    int lazy = 0;
    int ways = 1;          // bucket ways
    int scheme = 1;        // 2: chunk-wide window, per-piece tables seeded with the last occurrences in earlier pieces
    int pp = 8192;         // scheme 2: piece (tokenised independently, matches cut at its end)
    int tile = 1;          // scheme 2: 0 table only, 1 in-tile candidate preferred, 2 best of both
    int maxd = 32768;
    int lz4 = 0;
    int psamp = 1;         // scheme 2: the prepass (history for later pieces) records every psamp-th position only
    int bcap = 4;          // most bytes a match is extended backwards
    int dd = 4, sublanes = 16;   // tile >= 3: direct distances checked; tile >= 4: lanes per lookup/insert sub-step
    int tent = 0;          // scheme 2: table entries when not a power of two (multiply-shift range reduction)
    int nir = 0;           // scheme 2: positions inside a byte run (p-1..p+3 equal) are not inserted
    int bext = 0;          // backward extension of a selected match over the literals before it (inside the tile if 1, anywhere if 2)
};

static const uint16_t LBASE[29] = {3,4,5,6,7,8,9,10,11,13,15,17,19,23,27,31,35,43,51,59,67,83,99,115,131,163,195,227,258};
static const uint8_t LEXT[29] = {0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0};
static const uint16_t DBASE[30] = {1,2,3,4,5,7,9,13,17,25,33,49,65,97,129,193,257,385,513,769,1025,1537,2049,3073,4097,6145,8193,12289,16385,24577};
static const uint8_t DEXT[30] = {0,0,0,0,1,1,2,2,3,3,4,4,5,5,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13};
static int lsym(int len) { int s = 28; while (LBASE[s] > len) s--; return s; }
static int dsym(int d) { int s = 29; while (DBASE[s] > d) s--; return s; }

// optimal length-limited code lengths (simple: huffman then zlib-like fix-up by Kraft repair)
static void huff_lengths(const uint32_t *freq, int n, int maxbits, uint8_t *len)
{
    struct Node { uint64_t f; int l, r; };
    std::vector<Node> nodes; std::vector<int> alive;
    typedef std::pair<uint64_t, int> PI;
    std::priority_queue<PI, std::vector<PI>, std::greater<PI>> pq;
    for (int i = 0; i < n; i++) { len[i] = 0; if (freq[i]) { nodes.push_back({freq[i], -1 - i, 0}); pq.push({freq[i], (int)nodes.size() - 1}); } }
    if (nodes.empty()) return;
    if (nodes.size() == 1) { len[-1 - nodes[0].l] = 1; return; }
    while (pq.size() > 1) {
        PI a = pq.top(); pq.pop(); PI b = pq.top(); pq.pop();
        nodes.push_back({a.first + b.first, a.second, b.second}); pq.push({a.first + b.first, (int)nodes.size() - 1});
    }
    // depths
    std::vector<int> depth(nodes.size(), 0);
    for (int i = (int)nodes.size() - 1; i >= 0; i--) {
        if (nodes[i].l >= 0 || nodes[i].r > 0 || (nodes[i].l >= 0)) {}
        if (nodes[i].l >= 0 || nodes[i].l < 0) {
            if (!(nodes[i].l < 0 && nodes[i].r == 0)) { depth[nodes[i].l] = depth[i] + 1; depth[nodes[i].r] = depth[i] + 1; }
        }
    }
    for (size_t i = 0; i < nodes.size(); i++) if (nodes[i].l < 0 && nodes[i].r == 0) len[-1 - nodes[i].l] = (uint8_t)depth[i];
    // limit
    int over = 0; for (int i = 0; i < n; i++) if (len[i] > maxbits) { len[i] = (uint8_t)maxbits; over = 1; }
    if (over) {
        // Kraft repair: K = sum 2^(maxbits-len) must be <= 2^maxbits
        long K = 0; for (int i = 0; i < n; i++) if (len[i]) K += 1L << (maxbits - len[i]);
        while (K > (1L << maxbits)) {
            // lengthen the least frequent symbol with len < maxbits
            int best = -1; for (int i = 0; i < n; i++) if (len[i] && len[i] < maxbits && (best < 0 || freq[i] < freq[best] || (freq[i] == freq[best] && len[i] > len[best]))) best = i;
            K -= 1L << (maxbits - len[best] - 1); len[best]++;
        }
    }
}

static long dyn_header_bits(const uint8_t *ll, const uint8_t *dl, int *out_hlit = 0)
{
    int hlit = 286; while (hlit > 257 && ll[hlit - 1] == 0) hlit--;
    int hdist = 30; while (hdist > 1 && dl[hdist - 1] == 0) hdist--;
    std::vector<uint8_t> seq(ll, ll + hlit); seq.insert(seq.end(), dl, dl + hdist);
    uint32_t cf[19] = {0}; std::vector<std::pair<int,int>> items; // sym, extra bits count
    size_t i = 0;
    while (i < seq.size()) {
        size_t j = i; while (j < seq.size() && seq[j] == seq[i]) j++;
        size_t run = j - i; int v = seq[i];
        if (v == 0) {
            while (run >= 11) { size_t r = std::min<size_t>(run, 138); cf[18]++; items.push_back({18,7}); run -= r; }
            if (run >= 3) { cf[17]++; items.push_back({17,3}); run = 0; }
            while (run--) { cf[0]++; items.push_back({0,0}); }
        } else {
            cf[v]++; items.push_back({v,0}); run--;
            while (run >= 3) { size_t r = std::min<size_t>(run, 6); cf[16]++; items.push_back({16,2}); run -= r; }
            while (run--) { cf[v]++; items.push_back({v,0}); }
        }
        i = j;
    }
    uint8_t cl[19]; huff_lengths(cf, 19, 7, cl);
    static const uint8_t ORDER[19] = {16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15};
    int hclen = 19; while (hclen > 4 && cl[ORDER[hclen - 1]] == 0) hclen--;
    long bits = 5 + 5 + 4 + 3 * hclen;
    for (auto &it : items) bits += cl[it.first] + it.second;
    return bits;
}

struct Tok { int len, dist; uint8_t lit; };

static long encode_block_bits(const std::vector<Tok> &toks, int nbytes, long *dynb = 0, long *fixb = 0)
{
    uint32_t lf[286] = {0}, df[30] = {0};
    long extra = 0;
    for (auto &t : toks) {
        if (t.len == 0) lf[t.lit]++;
        else { int ls = lsym(t.len), ds = dsym(t.dist); lf[257 + ls]++; df[ds]++; extra += LEXT[ls] + DEXT[ds]; }
    }
    lf[256] = 1;
    uint8_t ll[286], dl[30]; huff_lengths(lf, 286, 15, ll); huff_lengths(df, 30, 15, dl);
    long dyn = 3 + dyn_header_bits(ll, dl) + extra;
    for (int i = 0; i < 286; i++) dyn += (long)lf[i] * ll[i];
    for (int i = 0; i < 30; i++) dyn += (long)df[i] * dl[i];
    long fix = 3 + extra;
    for (int i = 0; i < 286; i++) fix += (long)lf[i] * (i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8);
    for (int i = 0; i < 30; i++) fix += (long)df[i] * 5;
    long stored = 3 + 7 + 32 + 8L * nbytes;   // approx (alignment up to 7)
    if (dynb) *dynb = dyn; if (fixb) *fixb = fix;
    return std::min(dyn, std::min(fix, stored));
}

static inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint32_t hashf(const uint8_t *p, const Params &P)
{
    uint32_t v = rd32(p); if (P.hbytes == 3) v &= 0xffffff;
    if (P.tent) return (uint32_t)(((uint64_t)(v * 2654435761u) * (uint32_t)P.tent) >> 32);
    return (v * 2654435761u) >> (32 - P.hb);
}
static inline int mlen(const uint8_t *a, const uint8_t *b, int maxl) { int l = 0; while (l < maxl && a[l] == b[l]) l++; return l; }

// returns total bits for the piece (excluding final-block padding)
static long model_piece(const uint8_t *src, int n, const Params &P, long *ntok_out)
{
    std::vector<int> L(n, 0), D(n, 0);
    std::vector<int32_t> T((size_t)P.ways << P.hb, -1);
    for (int w0 = 0; w0 < n; w0 += P.W) {
        int w1 = std::min(n, w0 + P.W);
        for (int p = w0; p < w1; p++) {
            int maxl = std::min(258, n - p); int bl = 0, bd = 0;
            if (maxl >= P.minm && p + 4 <= n) {
                uint32_t h = hashf(src + p, P);
                for (int wy = 0; wy < P.ways; wy++) {
                    int c = T[(size_t)h * P.ways + wy];
                    if (c >= 0 && p - c <= 32768) { int l = mlen(src + c, src + p, maxl); if (l >= P.minm && l > bl) { bl = l; bd = p - c; } }
                }
                if (P.warp) {   // nearest lower lane in same warp with same hash
                    int base = p & ~31;
                    for (int q = p - 1; q >= base && q >= w0; q--) if (q + 4 <= n && hashf(src + q, P) == h) {
                        int l = mlen(src + q, src + p, maxl); if (l >= P.minm && l > bl) { bl = l; bd = p - q; } break; }
                }
            }
            if (P.rle && p >= 1 && maxl >= 3) { int l = mlen(src + p - 1, src + p, maxl); if (l >= 3 && l > bl) { bl = l; bd = 1; } }
            L[p] = bl; D[p] = bd;
        }
        // insert
        if (P.winner) { for (int p = w0; p < w1; p++) if (p + 4 <= n) { uint32_t h = hashf(src + p, P); if (P.ways > 1) { for (int wy = P.ways - 1; wy > 0; wy--) T[(size_t)h * P.ways + wy] = T[(size_t)h * P.ways + wy - 1]; } T[(size_t)h * P.ways] = p; } }
        else { for (int p = w1 - 1; p >= w0; p--) if (p + 4 <= n) T[(size_t)hashf(src + p, P) * P.ways] = p; }
    }
    long bits = 0, ntok = 0;
    for (int b0 = 0; b0 < n; b0 += P.sub) {
        int b1 = std::min(n, b0 + P.sub);
        std::vector<Tok> toks;
        int p = b0;
        while (p < b1) {
            int l = L[p];
            int lim = b1 - p; if (P.S) lim = std::min(lim, P.S - (p % P.S));
            if (l > lim) l = lim;
            if (P.lazy && l >= P.minm && l < 32 && p + 1 < b1) { int l2 = std::min(L[p + 1], lim - 1); if (l2 > l) l = 0; }
            if (l >= std::max(3, (D[p] == 1 ? 3 : P.minm)) ) { toks.push_back({l, D[p], 0}); p += l; }
            else { toks.push_back({0, 0, src[p]}); p++; }
        }
        ntok += toks.size();
        bits += encode_block_bits(toks, b1 - b0);
    }
    if (ntok_out) *ntok_out += ntok;
    return bits;
}


// scheme 2: the chunk (<= 64 KiB) sits whole in shared memory; piece k (pp bytes) has its own table of 2^hb entries, seeded
// before matching with the most recent occurrence of every hash in pieces < k; 32 positions per step look up, then insert.
static long lz4_size(const std::vector<Tok> &toks, int n);
static long model_chunk2(const uint8_t *src, int n, const Params &P, long *ntok_out)
{
    std::vector<char> Sel(n + 1, 0);
    const int np = (n + P.pp - 1) / P.pp, TS = P.tent ? P.tent : 1 << P.hb, WY = P.ways;
    std::vector<int> L(n, 0), D(n, 0);
    auto ins = [&](std::vector<int32_t> &T, uint32_t h, int p) { if (P.nir && p > 0 && rd32(src + p) == rd32(src + p - 1)) return; for (int w = WY - 1; w > 0; w--) T[h * WY + w] = T[h * WY + w - 1]; T[h * WY] = p; };
    std::vector<int32_t> carry(TS * WY, -1);
    for (int k = 0; k < np; k++) {
        std::vector<int32_t> T = carry;
        for (int p = k * P.pp; p < std::min(n, (k + 1) * P.pp); p++) if (p + 4 <= n && (p % P.psamp) == 0) ins(carry, hashf(src + p, P), p);
        const int e = std::min(n, (k + 1) * P.pp);
        int entry = k * P.pp, anchor = k * P.pp;
        for (int w0 = k * P.pp; w0 < e; w0 += 32) {
            const int w1 = std::min(e, w0 + 32);
            if (entry >= w1) continue;            // tile inside a running match: skipped, not indexed
            for (int p = w0; p < w1; p++) {
                const int maxl = std::min(258, e - p); int bl = 0, bd = 0;
                if (p + 4 <= n && maxl >= P.minm) {
                    const uint32_t h = hashf(src + p, P);
                    int ct = -1, cq = -1;
                    int lbest = -1; for (int w = 0; w < WY; w++) { const int c = T[h * WY + w]; if (c >= 0 && p - c <= P.maxd) { int l = mlen(src + c, src + p, maxl); if (l > lbest) { lbest = l; ct = c; } } }
                    if (P.tile == 1 || P.tile == 2) for (int q = p - 1; q >= w0; q--) if (rd32(src + q) == rd32(src + p)) { cq = q; break; }
                    if (P.tile >= 3) {
                        // direct short distances, then (tile >= 4) lanes of earlier sub-steps of `P.sub` lanes via the table
                        for (int d = 1; d <= P.dd && p - d >= 0; d++) if (rd32(src + p - d) == rd32(src + p)) { cq = p - d; break; }
                        if (cq < 0 && P.tile >= 4) { const int sb = w0 + ((p - w0) / P.sublanes) * P.sublanes; for (int q = sb - 1; q >= w0; q--) if (hashf(src + q, P) == h && !(P.nir && q > 0 && rd32(src + q) == rd32(src + q - 1))) { cq = q; break; } }
                    }
                    int lt = ct >= 0 ? mlen(src + ct, src + p, maxl) : 0, lq = cq >= 0 ? mlen(src + cq, src + p, maxl) : 0;
                    if (P.tile == 1 || P.tile >= 3) { if (cq >= 0) { bl = lq; bd = p - cq; } else { bl = lt; bd = p - ct; } }
                    else { if (lq >= lt && cq >= 0) { bl = lq; bd = p - cq; } else { bl = lt; bd = p - ct; } }
                    if (bl < P.minm) bl = 0;
                }
                L[p] = bl; D[p] = bd;
            }
            for (int p = w0; p < w1; p++) if (p + 4 <= n) ins(T, hashf(src + p, P), p);
            // greedy selection inside the tile
            int p = std::max(entry, w0);
            while (p < w1) {
                if (L[p]) {
                    int q = p;
                    if (P.bext) { const int lo = P.bext == 1 ? std::max(anchor, w0) : anchor; const int d = D[p];
                        while (q > lo && p - q < P.bcap && L[q] < 258 && q - 1 - d >= 0 && src[q - 1] == src[q - 1 - d]) { L[q - 1] = L[q] + 1; D[q - 1] = d; L[q] = 0; q--; } }
                    p = q + L[q]; anchor = p;
                } else p++;
            }
            entry = p;
        }
    }
    std::vector<Tok> toks;
    for (int k = 0; k < np; k++) {
        const int e = std::min(n, (k + 1) * P.pp);
        int p = k * P.pp;
        while (p < e) { if (L[p]) { toks.push_back({L[p], D[p], 0}); p += L[p]; } else { toks.push_back({0, 0, src[p]}); p++; } }
    }
    if (ntok_out) *ntok_out += toks.size();
    if (P.lz4) return 8 * lz4_size(toks, n);
    return encode_block_bits(toks, n);
}
static long lz4_size(const std::vector<Tok> &toks, int n)
{
    // LZ4 block: token byte, literal length ext, literals, offset(2), match length ext; last 5 bytes literal, last match starts >= 12 before end
    std::vector<Tok> mt;
    for (auto &t : toks) { if (t.len && !mt.empty() && mt.back().len && mt.back().dist == t.dist) mt.back().len += t.len; else mt.push_back(t); }
    long sz = 0; int lit = 0, pos = 0;
    for (auto &t : mt) {
        if (t.len == 0 || t.len < 4 || pos + t.len > n - 5 || pos > n - 12) { int l = t.len ? t.len : 1; lit += l; pos += l; continue; }
        sz += 1 + lit + (lit >= 15 ? 1 + (lit - 15) / 255 : 0) + 2 + (t.len - 4 >= 15 ? 1 + (t.len - 4 - 15) / 255 : 0);
        lit = 0; pos += t.len;
    }
    sz += 1 + lit + (lit >= 15 ? 1 + (lit - 15) / 255 : 0);
    return sz + 4;
}

int main(int argc, char **argv)
{
    Params P; size_t total = 24u << 20; int kind = 1; const char *only = 0; const char *file = 0;
    for (int i = 1; i < argc; i++) {
        auto eq = strchr(argv[i], '='); if (!eq) continue; int v = atoi(eq + 1); std::string k(argv[i], eq - argv[i]);
        if (k == "piece") P.piece = v; else if (k == "sub") P.sub = v; else if (k == "W") P.W = v; else if (k == "hb") P.hb = v;
        else if (k == "hbytes") P.hbytes = v; else if (k == "minm") P.minm = v; else if (k == "S") P.S = v; else if (k == "rle") P.rle = v;
        else if (k == "warp") P.warp = v; else if (k == "winner") P.winner = v; else if (k == "mb") total = (size_t)v << 20; else if (k == "kind") kind = v;
        else if (k == "lazy") P.lazy = v; else if (k == "ways") P.ways = v; else if (k == "only") only = eq + 1;
        else if (k == "scheme") P.scheme = v; else if (k == "pp") P.pp = v; else if (k == "tile") P.tile = v; else if (k == "maxd") P.maxd = v;
        else if (k == "lz4") P.lz4 = v; else if (k == "bext") P.bext = v; else if (k == "dd") P.dd = v; else if (k == "bcap") P.bcap = v; else if (k == "psamp") P.psamp = v; else if (k == "sublanes") P.sublanes = v; else if (k == "tent") P.tent = v; else if (k == "nir") P.nir = v; else if (k == "file") file = eq + 1;
    }
    std::vector<uint8_t> buf;
    if (file) {
        FILE *f = fopen(file, "rb"); if (!f) { perror(file); return 1; }
        fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET); if ((size_t)sz < total) total = sz;
        buf.resize(total); if (fread(buf.data(), 1, total, f) != total) return 1; fclose(f); kind = 0;
    } else {
        void *h = dlopen("harness/libqzcorpus.so", RTLD_NOW); if (!h) { fprintf(stderr, "no corpus lib\n"); return 1; }
        auto fill = (int (*)(int, uint64_t, uint64_t, uint8_t *, size_t, int))dlsym(h, "qzcorpus_fill");
        buf.resize(total); fill(kind, kind ? 0x51CE51A : 1, 0, buf.data(), total, 8);
    }
    int (*lz4c)(const char *, char *, int, int) = 0;
    if (P.lz4) { void *h = dlopen("liblz4.so.1", RTLD_NOW); if (h) lz4c = (int (*)(const char *, char *, int, int))dlsym(h, "LZ4_compress_default"); if (!lz4c) { fprintf(stderr, "no liblz4\n"); return 1; } }
    const char *CYC = "TXBETXBTZTXR";
    long zl_tot = 0, my_tot = 0; long zl_c[256] = {0}, my_c[256] = {0}, nb_c[256] = {0}, ntok = 0;
    std::vector<uint8_t> tmp(80000);
    for (size_t off = 0; off < total; off += 65536) {
        int n = (int)std::min<size_t>(65536, total - off); char cls = kind ? CYC[(off >> 20) % 12] : 'r';
        if (only && !strchr(only, cls)) continue;
        long zb;
        if (P.lz4) zb = lz4c((const char *)buf.data() + off, (char *)tmp.data(), n, (int)tmp.size()) + 4;
        else {
            z_stream z; memset(&z, 0, sizeof z); deflateInit2(&z, 1, Z_DEFLATED, -15, 9, 0);
            z.next_in = buf.data() + off; z.avail_in = n; z.next_out = tmp.data(); z.avail_out = tmp.size(); deflate(&z, Z_FINISH);
            zb = z.total_out; deflateEnd(&z);
        }
        long mb = 0;
        if (P.scheme == 2) mb = (model_chunk2(buf.data() + off, n, P, &ntok) + 7) / 8;
        else for (int q = 0; q < n; q += P.piece) mb += (model_piece(buf.data() + off + q, std::min(P.piece, n - q), P, &ntok) + 7) / 8 + (q + P.piece < n ? 5 : 0);
        zl_tot += zb; my_tot += mb; zl_c[(int)cls] += zb; my_c[(int)cls] += mb; nb_c[(int)cls] += n;
    }
    if (kind) for (int c = 0; c < 256; c++) if (nb_c[c]) printf("  %c ref %.4f model %.4f  rel %+.2f%%\n", c, (double)zl_c[c] / nb_c[c], (double)my_c[c] / nb_c[c], 100.0 * ((double)my_c[c] / zl_c[c] - 1));
    printf("TOTAL ref %ld (%.4f) model %ld (%.4f) rel %+.2f%%  tokens/byte %.3f\n", zl_tot, (double)zl_tot / total, my_tot, (double)my_tot / total, 100.0 * ((double)my_tot / zl_tot - 1), (double)ntok / total);
    return 0;
}
