// tests/cpu/hd_selftest.cpp -- runs the host+device scalar code of the codec (qz_huffman.h,
// qz_inflate.h, qz_deflate_tables.h, qz_crc32.h, qz_xxh32.h) on the CPU against zlib.
// The CUDA kernels execute these exact functions; the warp-parallel parts (sorting, scans,
// ballots) are emulated here by trivially serial code with the same semantics.
// Build+run: see tests/test_hd_core.py
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include <zlib.h>
#include "../../qatzip_b200/csrc/qz_huffman.h"
#include "../../qatzip_b200/csrc/qz_inflate.h"
#include "../../qatzip_b200/csrc/qz_crc32.h"
#include "../../qatzip_b200/csrc/qz_xxh32.h"
#include "../../qatzip_b200/csrc/qz_adler32.h"

static uint64_t rng_state = 88172645463325252ull;
static uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t)(rng_state >> 11); }

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #c); exit(1); } } while (0)

// ---- serial emulation of one piece of the compress kernel (same token semantics) ----
struct Tok { uint32_t t; };
static std::vector<uint8_t> compress_piece(const uint8_t *src, uint32_t n, bool bfinal, int force_type, int HB)
{
    std::vector<uint8_t> in(src, src + n); in.resize(n + 64, 0);
    std::vector<uint16_t> table(1u << HB, 0xffff);
    std::vector<uint32_t> toks; uint32_t hist[320] = {0}, extra = 0;
    auto ld32 = [&](uint32_t o) { uint32_t v; memcpy(&v, &in[o], 4); return v; };
    uint32_t entry = 0;
    for (uint32_t base = 0; base < n; base += 32) {
        uint32_t L[32], cand[32];
        for (uint32_t lane = 0; lane < 32; lane++) {
            uint32_t p = base + lane; L[lane] = 0; cand[lane] = 0xffff;
            if (p + 4 <= n) cand[lane] = table[(ld32(p) * 2654435761u) >> (32 - HB)];
        }
        for (uint32_t lane = 0; lane < 32; lane++) { uint32_t p = base + lane; if (p + 4 <= n) table[(ld32(p) * 2654435761u) >> (32 - HB)] = (uint16_t)p; }
        for (uint32_t lane = 0; lane < 32; lane++) {
            uint32_t p = base + lane; if (p >= n) continue;
            uint32_t maxl = std::min(258u, n - p);
            if (cand[lane] != 0xffff && ld32(cand[lane]) == ld32(p)) { uint32_t l = 4; while (l < maxl && in[cand[lane] + l] == in[p + l]) l++; L[lane] = std::min(l, maxl); }
        }
        uint32_t cur = entry;
        if (cur >= 32) { entry = cur - 32; continue; }
        while (cur < 32 && base + cur < n) {
            uint32_t p = base + cur;
            if (L[cur] >= 4) {
                uint32_t dist = p - cand[cur], ls, le, lv, ds, de, dv;
                qz_len_code(L[cur], &ls, &le, &lv); qz_dist_code(dist, &ds, &de, &dv);
                hist[ls]++; hist[288 + ds]++; extra += le + de;
                toks.push_back(0x80000000u | ((L[cur] - 3) << 16) | (dist - 1));
                cur += L[cur];
            } else { hist[in[p]]++; toks.push_back(in[p]); cur++; }
        }
        entry = cur >= 32 ? cur - 32 : 0;
    }
    hist[256] = 1;
    qz_huff_force_two(hist, QZ_NUM_LL); qz_huff_force_two(hist + 288, QZ_NUM_D);
    uint8_t ll_len[288] = {0}, d_len[32] = {0};
    uint32_t keys[512]; uint16_t ids[288]; int nk = 0;
    for (int s = 0; s < QZ_NUM_LL; s++) if (hist[s]) keys[nk++] = QZ_HUFF_KEY(hist[s], s);
    std::sort(keys, keys + nk); qz_huff_lengths_from_sorted(keys, ids, nk, 15, ll_len);
    nk = 0; for (int s = 0; s < QZ_NUM_D; s++) if (hist[288 + s]) keys[nk++] = QZ_HUFF_KEY(hist[288 + s], s);
    std::sort(keys, keys + nk); qz_huff_lengths_from_sorted(keys, ids, nk, 15, d_len);
    // Kraft check
    { double k = 0; for (int s = 0; s < 286; s++) if (ll_len[s]) { CHECK(ll_len[s] <= 15); k += 1.0 / (1u << ll_len[s]); } CHECK(k <= 1.0 + 1e-12 && k > 0.999999); }
    QzDynHeader hdr; qz_dyn_header_plan(ll_len, d_len, &hdr);
    uint32_t dynb = hdr.bits + extra, fixb = 3 + extra;
    for (int s = 0; s < 286; s++) { dynb += hist[s] * ll_len[s]; fixb += hist[s] * qz_fixed_ll_len(s); }
    for (int s = 0; s < 30; s++) { dynb += hist[288 + s] * d_len[s]; fixb += hist[288 + s] * 5; }
    uint32_t storedb = (5 + n) * 8;
    int btype = (dynb <= fixb && dynb < storedb) ? 2 : (fixb < storedb ? 1 : 0);
    if (n == 0) btype = 1;
    if (force_type >= 0) btype = force_type;
    std::vector<uint32_t> words((n + 1024) / 2 + 256, 0);
    std::vector<uint8_t> out;
    if (btype == 0) {
        out.push_back(bfinal ? 1 : 0); out.push_back((uint8_t)n); out.push_back((uint8_t)(n >> 8)); out.push_back((uint8_t)~n); out.push_back((uint8_t)(~n >> 8));
        out.insert(out.end(), src, src + n);
        return out;
    }
    uint32_t codes[320];
    if (btype == 1) { for (int s = 0; s < 288; s++) ll_len[s] = (uint8_t)qz_fixed_ll_len(s); for (int s = 0; s < 32; s++) d_len[s] = 5; }
    qz_huff_codes(ll_len, 288, codes); qz_huff_codes(d_len, btype == 1 ? 32 : 30, codes + 288);
    QzBitWriter bw; qz_bw_init(&bw, words.data());
    if (btype == 2) qz_dyn_header_write(&bw, &hdr, bfinal); else qz_bw_put(&bw, (bfinal ? 1u : 0u) | 2u, 3);
    if (btype == 2) CHECK(qz_bw_bitpos(&bw) == hdr.bits);
    uint32_t start_bits = qz_bw_bitpos(&bw);
    for (uint32_t t : toks) {
        if (t & 0x80000000u) {
            uint32_t ls, le, lv, ds, de, dv;
            qz_len_code(((t >> 16) & 0xff) + 3, &ls, &le, &lv); qz_dist_code((t & 0xffff) + 1, &ds, &de, &dv);
            qz_bw_put(&bw, codes[ls] & 0xffff, codes[ls] >> 16); if (le) qz_bw_put(&bw, lv, le);
            qz_bw_put(&bw, codes[288 + ds] & 0xffff, codes[288 + ds] >> 16); if (de) qz_bw_put(&bw, dv, de);
        } else qz_bw_put(&bw, codes[t] & 0xffff, codes[t] >> 16);
    }
    qz_bw_put(&bw, codes[256] & 0xffff, codes[256] >> 16);
    if (btype == 2) { uint32_t body = qz_bw_bitpos(&bw) - start_bits; CHECK(body <= dynb - hdr.bits + 2 && body + 32 >= dynb - hdr.bits); }
    if (!bfinal) { qz_bw_put(&bw, 0, 3); qz_bw_align_byte(&bw); qz_bw_put(&bw, 0, 16); qz_bw_put(&bw, 0xffff, 16); }
    uint32_t bytes = qz_bw_finish(&bw);
    out.assign((uint8_t *)words.data(), (uint8_t *)words.data() + bytes);
    return out;
}

// ---- serial driver of the inflate core (what the warp does around lane 0) ----
static int inflate_all(const uint8_t *src, uint32_t n, std::vector<uint8_t> &dst, uint32_t cap, uint32_t *consumed, bool raw_stop)
{
    static QzInflTables T;
    QzBitReader br; qz_br_init(&br, src, n);
    dst.assign(cap + 1, 0);
    uint32_t out = 0, bfinal = 0;
    while (!bfinal) {
        qz_br_refill(&br);
        if (raw_stop && qz_br_exhausted(&br)) break;
        bfinal = qz_br_bits(&br, 1); uint32_t type = qz_br_bits(&br, 2);
        if (type == 3) return -1;
        if (type == 0) {
            uint32_t drop = br.nacc & 7; br.acc >>= drop; br.nacc -= drop; qz_br_refill(&br);
            uint32_t len = qz_br_bits(&br, 16), nlen = qz_br_bits(&br, 16), start = qz_br_consumed(&br);
            if ((len ^ 0xffffu) != nlen) return -1;
            if (start + len > br.n) return -3;
            if (out + len > cap) return -2;
            memcpy(&dst[out], src + start, len); out += len;
            qz_br_seek(&br, start + len);
            continue;
        }
        uint32_t hlit = 288, hdist = 30;
        if (type == 1) qz_inflate_fixed_lens(&T); else if (qz_inflate_read_dynamic(&br, &T, &hlit, &hdist)) return -1;
        if (qz_infl_prepare(T.lens, hlit, T.ll_count, T.ll_first, T.ll_offs, T.ll_sorted) < 0) return -1;
        if (qz_infl_prepare(T.lens + hlit, hdist, T.d_count, T.d_first, T.d_offs, T.d_sorted) < 0) return -1;
        memset(T.ll_lut, 0, sizeof T.ll_lut); memset(T.d_lut, 0, sizeof T.d_lut);
        for (int lane = 0; lane < 32; lane++) {   // emulate the 32-lane fill
            qz_infl_fill_lut(T.lens, T.ll_count, T.ll_first, T.ll_offs, T.ll_sorted, T.ll_lut, QZ_LL_LUT_BITS, 0, lane, 32);
            qz_infl_fill_lut(T.lens + hlit, T.d_count, T.d_first, T.d_offs, T.d_sorted, T.d_lut, QZ_D_LUT_BITS, 1, lane, 32);
        }
        for (;;) {   // the kernel's batch loop: 32 tokens decoded without touching the output, then placed
            uint32_t tok[QZ_INFL_BATCH], ntk = 0, pos = out;
            int ev = qz_inflate_tokens(&br, &T, tok, &ntk, &pos, cap);
            for (uint32_t k = 0; k < ntk; k++) {
                if (!qz_tok_is_literal(tok[k])) { uint32_t ml = qz_tok_len(tok[k]), md = qz_tok_dist(tok[k]);
                    for (uint32_t q = 0; q < ml; q++) dst[out + q] = dst[out - md + (md >= ml ? q : q % md)]; out += ml; }
                else dst[out++] = (uint8_t)qz_tok_byte(tok[k]);
            }
            if (out != pos) return -9;
            if (ev == QZI_END_BLOCK) break;
            if (ev == QZI_ERR_DATA) return -1;
            if (ev == QZI_ERR_FULL) return -2;
            if (ev == QZI_ERR_TRUNC) return -3;
        }
    }
    if (qz_br_overrun(&br)) return -3;
    *consumed = qz_br_consumed(&br);
    dst.resize(out);
    return 0;
}

static std::vector<uint8_t> make_data(int kind, uint32_t n)
{
    std::vector<uint8_t> d(n);
    switch (kind) {
    case 0: for (auto &b : d) b = (uint8_t)rnd(); break;                                   // random
    case 1: for (auto &b : d) b = 0; break;                                                // zeros
    case 2: { const char *w[] = {"the ", "quick ", "brown ", "fox ", "jumps ", "over ", "lazy ", "dog. "}; uint32_t p = 0; while (p < n) { const char *s = w[rnd() % 8]; while (*s && p < n) d[p++] = *s++; } break; }
    case 3: for (uint32_t i = 0; i < n; i++) d[i] = (uint8_t)(i % 251 < 200 ? 'a' + (rnd() % 4) : rnd()); break;
    case 4: { uint32_t p = 0; while (p < n) { uint32_t run = rnd() % 100; uint8_t c = (uint8_t)(rnd() % 65 + 90); while (run-- && p < n) d[p++] = c; } break; }
    default: for (uint32_t i = 0; i < n; i++) d[i] = (uint8_t)((i * 7) ^ (i >> 5)); break;
    }
    return d;
}

int main()
{
    // 1. symbol arithmetic against the RFC tables
    static const uint16_t LBASE[29] = {3,4,5,6,7,8,9,10,11,13,15,17,19,23,27,31,35,43,51,59,67,83,99,115,131,163,195,227,258};
    static const uint8_t LEXT[29] = {0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0};
    static const uint16_t DBASE[30] = {1,2,3,4,5,7,9,13,17,25,33,49,65,97,129,193,257,385,513,769,1025,1537,2049,3073,4097,6145,8193,12289,16385,24577};
    static const uint8_t DEXT[30] = {0,0,0,0,1,1,2,2,3,3,4,4,5,5,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13};
    for (uint32_t s = 0; s < 29; s++) { uint32_t e, b = qz_len_base(s, &e); CHECK(b == LBASE[s] && e == LEXT[s]); }
    for (uint32_t s = 0; s < 30; s++) { uint32_t e, b = qz_dist_base(s, &e); CHECK(b == DBASE[s] && e == DEXT[s]); }
    for (uint32_t len = 3; len <= 258; len++) { uint32_t s, e, v; qz_len_code(len, &s, &e, &v); CHECK(s >= 257 && s <= 285 && LEXT[s - 257] == e && LBASE[s - 257] + v == len && (len == 258 ? s == 285 : true)); }
    for (uint32_t d = 1; d <= 32768; d++) { uint32_t s, e, v; qz_dist_code(d, &s, &e, &v); CHECK(s < 30 && DEXT[s] == e && DBASE[s] + v == d && v < (1u << e) + (e == 0)); }
    // 2. CRC-32 / combine / xxh32
    { std::vector<uint8_t> d = make_data(0, 100000); uint32_t tab[256]; for (int i = 0; i < 256; i++) tab[i] = qz_crc_table_entry(i);
      uint32_t c = 0xffffffffu; for (uint8_t b : d) c = tab[(c ^ b) & 0xff] ^ (c >> 8); c = ~c; CHECK(c == crc32(0, d.data(), d.size()));
      for (int t = 0; t < 50; t++) { uint32_t cut = rnd() % d.size(); uint32_t a = crc32(0, d.data(), cut), b = crc32(0, d.data() + cut, d.size() - cut); CHECK(qz_crc32_combine(a, b, d.size() - cut) == c); }
      CHECK(qz_xxh32((const uint8_t *)"", 0, 0) == 0x02CC5D05u);
      CHECK(qz_xxh32((const uint8_t *)"a", 1, 0) == 0x550D7456u);
      CHECK(qz_xxh32((const uint8_t *)"Nobody inspects the spammish repetition", 39, 0) == 0xE2293B2Fu);
      // Adler-32: whole-buffer, pairwise combine, and the warp's right-aligned 32-strip tree
      CHECK(qz_adler32(d.data(), 0) == 1u && qz_adler32(d.data(), d.size()) == adler32(1, d.data(), (uInt)d.size()));
      for (int t = 0; t < 50; t++) { uint32_t cut = rnd() % d.size(); uint32_t a = adler32(1, d.data(), cut), b = adler32(1, d.data() + cut, (uInt)(d.size() - cut));
                                     CHECK(qz_adler32_combine(a, b, d.size() - cut) == adler32(1, d.data(), (uInt)d.size())); }
      for (uint32_t n : {0u, 1u, 31u, 32u, 33u, 259u, 8192u, 8191u, 65536u, 99999u}) {
          const uint32_t S = n ? (n + 31) / 32 : 1; uint32_t s1[32], s2[32];
          for (int lane = 0; lane < 32; lane++) { int hi = (int)n - (31 - lane) * (int)S, lo = hi - (int)S; if (lo < 0) lo = 0; if (hi < lo) hi = lo;
                                                  qz_adler_block(d.data() + lo, (uint32_t)(hi - lo), &s1[lane], &s2[lane]); }
          for (int lv = 0; lv < 5; lv++) for (int lane = 0; lane < 32; lane += (2 << lv)) qz_adler_join(&s1[lane], &s2[lane], s1[lane + (1 << lv)], s2[lane + (1 << lv)], (uint64_t)S << lv);
          CHECK(qz_adler_finish(s1[0], s2[0], n) == adler32(1, d.data(), n));
      } }
    // 3. huffman length limiting on adversarial (fibonacci) frequencies
    { uint32_t f[40]; f[0] = 1; f[1] = 1; for (int i = 2; i < 40; i++) f[i] = f[i - 1] + f[i - 2] > 30000 ? 30000 : f[i - 1] + f[i - 2];
      for (int n = 2; n <= 40; n++) { uint32_t keys[64]; uint16_t ids[64]; uint8_t len[64] = {0}; for (int i = 0; i < n; i++) keys[i] = QZ_HUFF_KEY(f[i], i); std::sort(keys, keys + n);
        for (int cap = 7; cap <= 15; cap += 8) { if ((1 << cap) < n) continue; uint32_t k2[64]; memcpy(k2, keys, sizeof keys); memset(len, 0, sizeof len); qz_huff_lengths_from_sorted(k2, ids, n, cap, len);
          double k = 0; for (int i = 0; i < n; i++) { CHECK(len[i] >= 1 && len[i] <= cap); k += 1.0 / (1u << len[i]); } CHECK(k <= 1.0 + 1e-12); } } }
    // 4. compress emulation -> zlib inflate, and zlib deflate -> our inflate core
    uint32_t sizes[] = {0, 1, 2, 3, 4, 5, 31, 32, 33, 63, 64, 65, 100, 255, 256, 257, 1000, 4095, 4096, 8191, 8192, 16384};
    long total_in = 0, total_out = 0;
    for (int kind = 0; kind < 6; kind++) for (uint32_t n : sizes) for (int bfinal = 0; bfinal < 2; bfinal++) for (int ft = -1; ft <= 2; ft++) for (int HB = 11; HB <= 12; HB++) {
        std::vector<uint8_t> d = make_data(kind, n);
        std::vector<uint8_t> c = compress_piece(d.data(), n, bfinal, ft, HB);
        if (ft < 0 && HB == 11) { total_in += n; total_out += c.size(); }
        // zlib raw inflate
        std::vector<uint8_t> o(n + 16);
        z_stream z; memset(&z, 0, sizeof z); CHECK(inflateInit2(&z, -15) == Z_OK);
        z.next_in = c.data(); z.avail_in = (uInt)c.size(); z.next_out = o.data(); z.avail_out = (uInt)o.size();
        int r = inflate(&z, Z_SYNC_FLUSH);
        CHECK(bfinal ? r == Z_STREAM_END : (r == Z_OK || r == Z_BUF_ERROR));
        CHECK(z.total_out == n && memcmp(o.data(), d.data(), n) == 0);
        CHECK(z.avail_in == 0);
        inflateEnd(&z);
        // our inflate core on our own stream
        std::vector<uint8_t> o2; uint32_t used = 0;
        CHECK(inflate_all(c.data(), (uint32_t)c.size(), o2, n, &used, !bfinal) == 0);
        CHECK(o2.size() == n && memcmp(o2.data(), d.data(), n) == 0 && used == c.size());
    }
    for (int kind = 0; kind < 6; kind++) for (uint32_t n : {0u, 1u, 100u, 5000u, 70000u, 300000u}) for (int level : {1, 6, 9}) for (int strat : {Z_DEFAULT_STRATEGY, Z_FIXED, Z_HUFFMAN_ONLY}) {
        std::vector<uint8_t> d = make_data(kind, n), c(n + n / 8 + 256);
        z_stream z; memset(&z, 0, sizeof z); CHECK(deflateInit2(&z, level, Z_DEFLATED, -15, 9, strat) == Z_OK);
        z.next_in = d.data(); z.avail_in = n; z.next_out = c.data(); z.avail_out = (uInt)c.size();
        // flush in the middle like the reference's software path (Z_FULL_FLUSH per hw_buff_sz)
        if (n > 65536) { z.avail_in = 65536; CHECK(deflate(&z, Z_FULL_FLUSH) == Z_OK); z.avail_in = n - 65536; }
        CHECK(deflate(&z, Z_FINISH) == Z_STREAM_END); c.resize(z.total_out); deflateEnd(&z);
        std::vector<uint8_t> o; uint32_t used = 0;
        CHECK(inflate_all(c.data(), (uint32_t)c.size(), o, n, &used, false) == 0);
        CHECK(o.size() == n && (n == 0 || memcmp(o.data(), d.data(), n) == 0) && used == c.size());
        if (n > 10) { std::vector<uint8_t> o3; CHECK(inflate_all(c.data(), (uint32_t)c.size() - 3, o3, n, &used, false) != 0);       // truncated input
                      CHECK(inflate_all(c.data(), (uint32_t)c.size(), o3, n - 1, &used, false) == -2); }                               // output too small
    }
    printf("hd_selftest ok (emulated piece ratio on mixed toy data: %.4f)\n", (double)total_out / (double)total_in);
    return 0;
}
