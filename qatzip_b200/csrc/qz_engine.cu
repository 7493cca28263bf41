/* qz_engine.cu -- host-side chunk engine (see qz_engine.h).
 *
 * Reference loops replaced here:
 *   compress    doCompressIn  (src/qatzip.c:1483-1604)  chunk loop, staging copy, submit
 *               doCompressOut (src/qatzip.c:1610-1764)  poll, dest-space check, stitch, CRC combine
 *   decompress  doDecompressIn/Out (src/qatzip.c:2103-2404) + checkHeader (src/qatzip_utils.c:1232)
 * QAT's DMA rings become: pinned pages + cudaMemcpyAsync on per-slot streams, two slots in
 * flight so the H2D of batch k+1 and the D2H of batch k-1 overlap the kernels of batch k.
 * There is no polling thread and no sleep/back-off: completion is a CUDA event.
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <chrono>
#include <mutex>
#include <vector>
#include <thread>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <algorithm>
#include "qz_engine.h"
#include "qz_kernels.cuh"
#include "qz_crc32.h"
#include "qz_adler32.h"
#include "qz_xxh32.h"

extern "C" cudaError_t qzb_launch_deflate(const QzbCompressJob *job, int hb, int grid, int warps, int nbuf, cudaStream_t st);
extern "C" cudaError_t qzb_launch_deflate_window(const QzbCompressJob *job, int grid, int ways, cudaStream_t st);
extern "C" size_t qzb_deflate_window_tok_words(int grid);
extern "C" int qzb_deflate_window_max_tent(void);
extern "C" cudaError_t qzb_launch_frame(const QzbCompressJob *job, cudaStream_t st);
extern "C" cudaError_t qzb_launch_inflate(const QzbDecompressJob *job, int dpw, int grid, cudaStream_t st);
extern "C" size_t qzb_inflate_smem_bytes(int dpw);
extern "C" int qzb_inflate_cta_threads(int dpw);
extern "C" int qzb_inflate_cta_slots(int dpw);
extern "C" cudaError_t qzb_launch_gzip_scan(const uint8_t *p, uint64_t lo, uint64_t n, uint32_t *list, uint32_t cap, uint32_t *count, int grid, cudaStream_t st);
extern "C" cudaError_t qzb_launch_lz4_compress(const QzbCompressJob *job, int grid, int warps, cudaStream_t st);
extern "C" cudaError_t qzb_launch_lz4_decompress(const QzbDecompressJob *job, int grid, cudaStream_t st);
extern "C" size_t qzb_deflate_smem_bytes(int piece_log2, int hb, int warps, int nbuf);
extern "C" size_t qzb_deflate_window_smem_bytes(int tent);
extern "C" size_t qzb_lz4_smem_bytes(int piece_log2, int warps);
extern "C" cudaError_t qzb_launch_lz4_window(const QzbCompressJob *job, int grid, int nw, cudaStream_t st);
extern "C" size_t qzb_lz4_window_smem_bytes(int tent, int nw);
extern "C" int qzb_lz4_window_max_tent(int nw);
extern "C" size_t qzb_lz4_window_tok_words(int grid, int nw);

/* qatzip.h return codes used here (kept numeric so this file does not depend on the public header) */
enum { RC_OK = 0, RC_PARAMS = -1, RC_FAIL = -2, RC_BUF_ERROR = -3, RC_DATA_ERROR = -4 };

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { qzb_log_cuda(#call, e_, __LINE__); return RC_FAIL; } } while (0)
static void qzb_log_cuda(const char *what, cudaError_t e, int line)
{
    if (getenv("QZB200_DEBUG")) fprintf(stderr, "[qatzip_b200] %s failed at qz_engine.cu:%d: %s\n", what, line, cudaGetErrorString(e));
}

/* ------------------------------------------------------------------ runtime + tuning */
static std::once_flag g_rt_once;
static int g_ndev = 0;
extern "C" int qzb_runtime_devices(void)
{
    std::call_once(g_rt_once, [] {
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess) { n = 0; (void)cudaGetLastError(); }
        g_ndev = n;
    });
    return g_ndev;
}
/* most warps per CTA the deflate kernels were compiled for (their launch bound; qz_deflate.cu) */
extern "C" int qzb_deflate_max_warps(int group);
#ifndef QZB200_WINDOW_DEFAULT
#define QZB200_WINDOW_DEFAULT 1
#endif
static int env_int(const char *name, int dflt) { const char *v = getenv(name); return (v && *v) ? atoi(v) : dflt; }
extern "C" int qzb_runtime_default_device(void)
{
    int n = qzb_runtime_devices();
    if (n <= 0) return -1;
    int d = env_int("QZB200_DEVICE", -1);
    if (d < 0) d = env_int("LOCAL_RANK", 0);
    return d % n;
}
/* Devices one engine may spread a host-buffer compress call over (the reference interleaves its instances across the QAT
 * devices of a process, src/qatzip.c:795-808 and qzGrabInstance :363): QZB200_DEVICES = "all", a count, or a comma-separated
 * list; unset = the one default device (one process per GPU, as under torchrun). */
static std::vector<int> configured_devices(void)
{
    std::vector<int> d;
    const int n = qzb_runtime_devices();
    const char *v = getenv("QZB200_DEVICES");
    if (n <= 0) return d;
    if (!v || !*v) { d.push_back(qzb_runtime_default_device()); return d; }
    if (!strcmp(v, "all")) { for (int i = 0; i < n; i++) d.push_back(i); return d; }
    if (!strchr(v, ',')) { const int k = std::max(1, std::min(n, atoi(v))); const int first = qzb_runtime_default_device(); for (int i = 0; i < k; i++) d.push_back((first + i) % n); return d; }
    for (const char *p = v; *p;) { const int x = atoi(p); if (x >= 0 && x < n && std::find(d.begin(), d.end(), x) == d.end()) d.push_back(x); p = strchr(p, ','); if (!p) break; p++; }
    if (d.empty()) d.push_back(qzb_runtime_default_device());
    return d;
}
extern "C" int qzb_runtime_device_list(int *out, int cap)
{
    const std::vector<int> d = configured_devices();
    for (int i = 0; i < (int)d.size() && i < cap; i++) out[i] = d[i];
    return (int)d.size();
}
extern "C" void qzb_get_tuning(QzbTuning *t)
{
    t->piece_log2 = env_int("QZB200_PIECE_LOG2", 13);
    if (t->piece_log2 != 13 && t->piece_log2 != 14) t->piece_log2 = 13;
    t->hash_bits = env_int("QZB200_HASH_BITS", t->piece_log2 == 13 ? 11 : 12);
    if (t->piece_log2 == 13 && t->hash_bits != 11 && t->hash_bits != 12) t->hash_bits = 11;
    if (t->piece_log2 == 14 && t->hash_bits != 12 && t->hash_bits != 13) t->hash_bits = 12;
    t->warps_per_cta = env_int("QZB200_WARPS", 0);       /* 0 = default geometry */
    t->buffers_per_cta = env_int("QZB200_BUFFERS", 0);
    t->inflate_dpw = env_int("QZB200_INFLATE_DPW", 1);      /* members decoded at once by one warp: 1 (measured fastest on B200), 2, 4 or 8 */
    if (t->inflate_dpw != 2 && t->inflate_dpw != 4 && t->inflate_dpw != 8) t->inflate_dpw = 1;
    int mb = env_int("QZB200_BATCH_MB", 64);
    if (mb < 1) mb = 1;
    if (mb > 1024) mb = 1024;
    t->batch_bytes = (size_t)mb << 20;
    /* Decompress batches (compressed bytes per launch, host-memory calls).  A launch lasts at least as long as its longest
     * member takes one warp (a 256 KiB member: ~25 ms), so the bytes in flight -- four batches -- have to cover that time at
     * the kernel's rate.  Measured on mixed 4-256 KiB members, a 4 GiB call: 64 MiB batches 12.5 GB/s, 128 MiB 23-24,
     * 256 MiB 28.1, 512 MiB 28.6 (device buffers grow with the batch: 256 MiB in + up to 1 GiB out per slot, only for calls
     * that large) */
    int imb = env_int("QZB200_INFLATE_BATCH_MB", 256);
    if (imb < 1) imb = 1;
    if (imb > 1024) imb = 1024;
    t->inflate_batch_bytes = (size_t)imb << 20;
    int fmb = env_int("QZB200_FIRST_MB", 16);
    if (fmb < 1) fmb = 1;
    if (fmb > mb) fmb = mb;
    t->first_batch_bytes = (size_t)fmb << 20;
    t->taper = env_int("QZB200_TAPER", 0);
    /* deflate, hw_buff_sz >= 64 KiB: 1 = window kernel (64 KiB windows in shared memory, one block per window), 0 = one block per
     * 8 KiB piece with a private window everywhere */
    t->window = env_int("QZB200_WINDOW", QZB200_WINDOW_DEFAULT);
    t->window_tent = env_int("QZB200_WINDOW_TENT", 0);      /* entries of a matcher's hash table (2 bytes each, thirty tables); 0 = as many as fit */
    if (t->window_tent && (t->window_tent < 256 || t->window_tent > 8192)) t->window_tent = 0;
    t->lz4_warps = env_int("QZB200_LZ4_WARPS", 0);          /* LZ4 window kernel: 16 warps (tables of ~5200 entries) or 12 (~6900); 0 = by level */
    int wmb = env_int("QZB200_ZLIB_WINDOW_MB", 128);
    if (wmb < 1) wmb = 1;
    if (wmb > 2048) wmb = 2048;
    t->zlib_window_bytes = (size_t)wmb << 20;
}

/* ------------------------------------------------------------------ pinned registry */
static std::mutex g_pin_lock;
static std::map<uintptr_t, size_t> g_pinned;
extern "C" void *qzb_pinned_alloc(size_t sz)
{
    if (qzb_runtime_devices() <= 0) return NULL;
    void *p = NULL;
    if (cudaHostAlloc(&p, sz ? sz : 1, cudaHostAllocPortable) != cudaSuccess) { (void)cudaGetLastError(); return NULL; }
    std::lock_guard<std::mutex> g(g_pin_lock);
    g_pinned[(uintptr_t)p] = sz ? sz : 1;
    return p;
}
extern "C" int qzb_pinned_free(void *p)
{
    {
        std::lock_guard<std::mutex> g(g_pin_lock);
        auto it = g_pinned.find((uintptr_t)p);
        if (it == g_pinned.end()) return 0;
        g_pinned.erase(it);
    }
    cudaFreeHost(p);
    return 1;
}
extern "C" int qzb_pinned_contains(const void *p, size_t len)
{
    std::lock_guard<std::mutex> g(g_pin_lock);
    auto it = g_pinned.upper_bound((uintptr_t)p);
    if (it == g_pinned.begin()) return 0;
    --it;
    return (uintptr_t)p >= it->first && (uintptr_t)p + (len ? len : 1) <= it->first + it->second;
}

extern "C" void *qzb_device_alloc(int device, size_t n)
{
    void *p = NULL;
    if (cudaSetDevice(device) != cudaSuccess || cudaMalloc(&p, n ? n : 1) != cudaSuccess) { (void)cudaGetLastError(); return NULL; }
    return p;
}
extern "C" void qzb_device_free(int device, void *p) { if (p && cudaSetDevice(device) == cudaSuccess) cudaFree(p); }
extern "C" int qzb_device_copy(int device, void *dst, const void *src, size_t n, int to_device)
{
    if (cudaSetDevice(device) != cudaSuccess) return RC_FAIL;
    return cudaMemcpy(dst, src, n, to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost) == cudaSuccess ? RC_OK : RC_FAIL;
}

/* ------------------------------------------------------------------ buffers */
struct DevBuf {
    void *p = nullptr; size_t cap = 0;
    int ensure(size_t n) {
        if (n <= cap) return RC_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + (n >> 3) + 4096;
        if (cudaMalloc(&p, want) != cudaSuccess) { (void)cudaGetLastError(); if (cudaMalloc(&p, n) != cudaSuccess) { (void)cudaGetLastError(); p = nullptr; return RC_FAIL; } want = n; }
        cap = want; return RC_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct HostBuf {
    void *p = nullptr; size_t cap = 0;
    int ensure(size_t n) {
        if (n <= cap) return RC_OK;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = n + (n >> 3) + 4096;
        if (cudaHostAlloc(&p, want, cudaHostAllocPortable) != cudaSuccess) { (void)cudaGetLastError(); p = nullptr; return RC_FAIL; }
        cap = want; return RC_OK;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct Slot {
    int device = 0;
    cudaStream_t st = nullptr;
    cudaEvent_t ev_k0 = nullptr, ev_km = nullptr, ev_k1 = nullptr, ev_meta = nullptr, ev_done = nullptr, ev_h0 = nullptr, ev_d0 = nullptr, ev_d1 = nullptr;
    DevBuf d_in, d_slots, d_out, d_meta, d_tok, d_members, d_results;
    HostBuf h_meta, h_in, h_out, h_members, h_results;
    /* bookkeeping of the batch currently in flight */
    bool busy = false;
    uint64_t in_off = 0, in_len = 0;
    uint32_t nchunks = 0;
    size_t first_member = 0, nmembers = 0;
    uint64_t span_src = 0, span_len = 0, out_base = 0, out_len = 0;
};

struct QzbEngine {
    int device = 0;                       /* primary device: decompress, device-resident calls */
    std::vector<int> devices;             /* devices[0] == device; host-buffer compress calls deal their batches over all of them */
    int sm_count = 148;
    QzbTuning tune;
    static constexpr int NSLOT = 4;       /* batches in flight per device: copy-in queued, copy-in, compute, copy-out */
    static constexpr int MAX_DEV = 16;
    Slot slot[NSLOT * MAX_DEV];           /* slot[d * NSLOT + k]: k-th slot of devices[d] */
    cudaEvent_t ev_base = nullptr;        /* QZB200_TIMELINE=1: start of the host-buffer call, for per-batch timestamps on stderr */
    int timeline = 0;
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int slot_setup(Slot &s, int device)
{
    s.device = device;
    if (cudaSetDevice(device) != cudaSuccess) return RC_FAIL;
    if (cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking) != cudaSuccess) return RC_FAIL;
    cudaEventCreate(&s.ev_k0); cudaEventCreate(&s.ev_km); cudaEventCreate(&s.ev_k1); cudaEventCreate(&s.ev_h0); cudaEventCreate(&s.ev_d0); cudaEventCreate(&s.ev_d1);
    cudaEventCreateWithFlags(&s.ev_meta, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming);
    return RC_OK;
}
/* `device`: the engine's primary device.  With QZB200_DEVICES naming several, the engine also gets slots on the others
 * (in list order, starting behind the primary) for host-buffer compress calls. */
extern "C" QzbEngine *qzb_engine_create(int device)
{
    if (device < 0 || device >= qzb_runtime_devices()) return NULL;
    if (cudaSetDevice(device) != cudaSuccess) { (void)cudaGetLastError(); return NULL; }
    QzbEngine *e = new QzbEngine();
    e->device = device;
    e->devices.push_back(device);
    {
        const std::vector<int> all = configured_devices();
        const auto it = std::find(all.begin(), all.end(), device);
        if (it != all.end())
            for (size_t k = 1; k < all.size() && (int)e->devices.size() < QzbEngine::MAX_DEV; k++) e->devices.push_back(all[((it - all.begin()) + k) % all.size()]);
    }
    qzb_get_tuning(&e->tune);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) e->sm_count = prop.multiProcessorCount;
    e->timeline = env_int("QZB200_TIMELINE", 0);
    cudaEventCreate(&e->ev_base);
    for (size_t d = 0; d < e->devices.size(); d++)
        for (int k = 0; k < QzbEngine::NSLOT; k++)
            if (slot_setup(e->slot[d * QzbEngine::NSLOT + k], e->devices[d]) != RC_OK) { (void)cudaGetLastError(); qzb_engine_destroy(e); return NULL; }
    cudaSetDevice(device);
    return e;
}
extern "C" void qzb_engine_destroy(QzbEngine *e)
{
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->ev_base) cudaEventDestroy(e->ev_base);
    for (auto &s : e->slot) {
        if (!s.st) continue;
        cudaSetDevice(s.device);
        cudaStreamSynchronize(s.st); cudaStreamDestroy(s.st);
        if (s.ev_k0) cudaEventDestroy(s.ev_k0);
        if (s.ev_km) cudaEventDestroy(s.ev_km);
        if (s.ev_h0) cudaEventDestroy(s.ev_h0);
        if (s.ev_d0) cudaEventDestroy(s.ev_d0);
        if (s.ev_d1) cudaEventDestroy(s.ev_d1);
        if (s.ev_k1) cudaEventDestroy(s.ev_k1);
        if (s.ev_meta) cudaEventDestroy(s.ev_meta);
        if (s.ev_done) cudaEventDestroy(s.ev_done);
        s.d_in.release(); s.d_slots.release(); s.d_out.release(); s.d_meta.release(); s.d_tok.release();
        s.d_members.release(); s.d_results.release();
        s.h_meta.release(); s.h_in.release(); s.h_out.release(); s.h_members.release(); s.h_results.release();
    }
    delete e;
}
extern "C" int qzb_engine_device_count(const QzbEngine *e) { return e ? (int)e->devices.size() : 0; }
extern "C" int qzb_engine_primary_device(const QzbEngine *e) { return e ? e->device : -1; }

/* Whatever way an engine call ends, no slot is left with work in flight or marked busy: a later call on the session
 * would otherwise drain a stale batch into its own output (and copies into the caller's buffers would still be running). */
struct SlotGuard {
    QzbEngine *e;
    explicit SlotGuard(QzbEngine *e_) : e(e_) { settle(); }
    ~SlotGuard() { settle(); cudaSetDevice(e->device); }
    void settle() { for (auto &s : e->slot) if (s.st && s.busy) { cudaSetDevice(s.device); cudaStreamSynchronize(s.st); (void)cudaGetLastError(); s.busy = false; } }
};

/* ------------------------------------------------------------------ compress */
static uint32_t hdr_sz(int fmt) { return fmt == QZB_FMT_GZIP_EXT ? 24u : fmt == QZB_FMT_GZIP ? 10u : fmt == QZB_FMT_4B ? 4u : fmt == QZB_FMT_LZ4 ? 15u : fmt == QZB_FMT_ZLIB ? 2u : 0u; }
static uint32_t ftr_sz(int fmt) { return (fmt == QZB_FMT_GZIP_EXT || fmt == QZB_FMT_GZIP || fmt == QZB_FMT_LZ4) ? 8u : fmt == QZB_FMT_ZLIB ? 4u : 0u; }

struct MetaLayout { size_t piece_len, piece_crc, chunk_total, chunk_cksum, chunk_off, ticket, total; };
static MetaLayout meta_layout(uint32_t npieces, uint32_t nchunks)
{
    MetaLayout m; size_t o = 0;
    m.chunk_off = o; o += align_up((size_t)(nchunks + 1) * 8, 16);
    m.chunk_cksum = o; o += align_up((size_t)nchunks * 4, 16);
    m.chunk_total = o; o += align_up((size_t)nchunks * 4, 16);
    m.piece_len = o; o += align_up((size_t)npieces * 4, 16);
    m.piece_crc = o; o += align_up((size_t)npieces * 4, 16);
    m.ticket = o; o += 16;
    m.total = o;
    return m;
}

/* Enqueue one batch on a slot's stream: kernels + D2H of the per-chunk offsets and checksums.
 * d_src/d_dst are device pointers (the slot's own buffers, or the caller's for device-resident calls). */
static int enqueue_compress(QzbEngine *e, Slot &s, const QzbCompressCall *c, const uint8_t *d_src, uint64_t len,
                            uint8_t *d_dst, uint64_t d_dst_cap, int last, uint64_t *launches)
{
    const QzbTuning &t = e->tune;
    const uint32_t PIECE = 1u << t.piece_log2;
    CK(cudaSetDevice(s.device));
    QzbCompressJob job; memset(&job, 0, sizeof job);
    job.src = d_src; job.src_len = len; job.chunk_sz = c->chunk_sz; job.piece_log2 = (uint32_t)t.piece_log2;
    job.pieces_per_chunk = (c->chunk_sz + PIECE - 1) / PIECE;
    job.nchunks = len ? (uint32_t)((len + c->chunk_sz - 1) / c->chunk_sz) : 1u;
    const uint64_t last_len = len - (uint64_t)(job.nchunks - 1) * c->chunk_sz;
    const uint32_t last_pieces = last_len ? (uint32_t)((last_len + PIECE - 1) / PIECE) : 1u;
    job.npieces = (job.nchunks - 1) * job.pieces_per_chunk + last_pieces;
    job.fmt = c->fmt; job.last = last; job.static_huffman = c->static_huffman;
    job.slot_stride = PIECE + 64;
    /* launch geometry: fill the SMs with as many warps as shared memory allows */
    int warps = t.warps_per_cta, ctas_per_sm = 1;
    const size_t smem_cap = 227 * 1024;
    const bool lz4 = (c->fmt == QZB_FMT_LZ4);
    /* deflate: NW warps share NB piece buffers (NW ~ 2 NB, see qz_deflate.cu); LZ4 warps each own one */
    int nbuf = t.buffers_per_cta;
    size_t group_smem = 0;
    auto smem_for = [&](int w, int nb) { return lz4 ? qzb_lz4_smem_bytes(t.piece_log2, w) : qzb_deflate_smem_bytes(t.piece_log2, t.hash_bits, w, nb); };
    int lz4_nw = 0;
    if (lz4 && t.window && t.piece_log2 == 13 && len && job.pieces_per_chunk % 8 == 0) {
        /* LZ4 window kernel: one block per 64 KiB window, one CTA per SM, tables as large as the shared memory allows */
        const uint32_t wpc = job.pieces_per_chunk / 8;
        job.ngroups = (job.nchunks - 1) * wpc + (last_pieces + 7) / 8;
        /* levels below 6: sixteen matchers (+19 % throughput, +1.3 % output on the bench corpus); 6 and up: twelve with larger tables */
        lz4_nw = t.lz4_warps == 16 ? 16 : t.lz4_warps == 12 ? 12 : (c->level >= 6 ? 12 : 16);
        const int fit = qzb_lz4_window_max_tent(lz4_nw);
        job.tent = (uint32_t)(t.window_tent > 0 ? std::min(t.window_tent, fit) : fit);
        group_smem = qzb_lz4_window_smem_bytes((int)job.tent, lz4_nw);
        warps = lz4_nw; nbuf = 0;
    } else if (lz4) {
        if (warps <= 0 || warps > 16) warps = 16;
        while (warps > 1 && smem_for(warps, 0) + 2304 > smem_cap) warps--;
        nbuf = 0;
    } else if (t.window && t.piece_log2 == 13 && len && job.pieces_per_chunk % 8 == 0) {
        /* window kernel: CTAs of whole groups of 8 warps; as many units as the 227 KB hold, never more than groups */
        const uint32_t wpc = job.pieces_per_chunk / 8;
        job.ngroups = (job.nchunks - 1) * wpc + (last_pieces + 7) / 8;
        /* the thirty tables take what the window and the block coders leave of the 227 KB (QZB200_WINDOW_TENT: fewer entries) */
        const int fit = qzb_deflate_window_max_tent();
        job.tent = (uint32_t)(t.window_tent > 0 ? std::min(t.window_tent, fit) : fit);
        group_smem = qzb_deflate_window_smem_bytes((int)job.tent);
        warps = 32;
    } else {
        if (warps <= 0 || warps > qzb_deflate_max_warps(0)) { warps = 20; if (nbuf <= 0) nbuf = 17; }
        if (nbuf <= 0 || nbuf > warps) nbuf = (warps + 1) / 2;
        while (warps > 2 && smem_for(warps, nbuf) + 2304 > smem_cap) { warps -= 2; nbuf = std::min(nbuf, (warps + 1) / 2); }
    }
    ctas_per_sm = (int)std::max<size_t>(1, (smem_cap + 1024) / ((group_smem ? group_smem : smem_for(warps, nbuf)) + 3072));
    if (ctas_per_sm * warps > 48) ctas_per_sm = std::max(1, 48 / warps);
    int grid = e->sm_count * ctas_per_sm;
    /* small batches: one unit (group or piece) per CTA spreads them over the SMs -- a group alone on an SM finishes in a
     * fraction of the time it takes next to three others, and the surplus warps of each CTA find no ticket and leave */
    const int need = job.ngroups ? (int)job.ngroups : (int)job.npieces;
    if (grid > need) grid = std::max(1, need);

    if (s.d_slots.ensure((size_t)job.npieces * job.slot_stride + 64) != RC_OK) return RC_FAIL;
    MetaLayout ml = meta_layout(job.npieces, job.nchunks);
    if (s.d_meta.ensure(ml.total) != RC_OK || s.h_meta.ensure(ml.piece_len) != RC_OK) return RC_FAIL;
    if (s.d_tok.ensure((lz4_nw ? qzb_lz4_window_tok_words(grid, lz4_nw) : job.ngroups ? qzb_deflate_window_tok_words(grid) : (size_t)grid * warps * QZB_TOK_STRIDE(PIECE)) * 4) != RC_OK) return RC_FAIL;
    uint8_t *dm = (uint8_t *)s.d_meta.p;
    job.slots = (uint8_t *)s.d_slots.p;
    job.piece_len = (uint32_t *)(dm + ml.piece_len); job.piece_crc = (uint32_t *)(dm + ml.piece_crc);
    job.chunk_total = (uint32_t *)(dm + ml.chunk_total); job.chunk_cksum = (uint32_t *)(dm + ml.chunk_cksum);
    job.chunk_off = (uint64_t *)(dm + ml.chunk_off); job.ticket = (uint32_t *)(dm + ml.ticket);
    job.tok_scratch = (uint32_t *)s.d_tok.p;
    job.dst = d_dst; job.dst_cap = d_dst_cap;

    CK(cudaMemsetAsync(job.ticket, 0, 16, s.st));
    CK(cudaEventRecord(s.ev_k0, s.st));
    if (lz4_nw) CK(qzb_launch_lz4_window(&job, grid, lz4_nw, s.st));
    else if (lz4) CK(qzb_launch_lz4_compress(&job, grid, warps, s.st));
    else if (job.ngroups) CK(qzb_launch_deflate_window(&job, grid, c->level >= 6 ? 2 : 1, s.st));      /* levels 6 and up: two-entry buckets (QAT 2.0 maps 1-5 / 6-8 / 9-12 to three search depths, reference README.md:133-148) */
    else CK(qzb_launch_deflate(&job, t.hash_bits, grid, warps, nbuf, s.st));
    CK(cudaEventRecord(s.ev_km, s.st));
    CK(qzb_launch_frame(&job, s.st));
    CK(cudaEventRecord(s.ev_k1, s.st));
    *launches += lz4 ? 5 : 4;
    /* chunk_off + chunk_cksum are adjacent at the start of the meta block */
    CK(cudaMemcpyAsync(s.h_meta.p, dm, ml.chunk_total, cudaMemcpyDeviceToHost, s.st));
    CK(cudaEventRecord(s.ev_meta, s.st));
    s.nchunks = job.nchunks;
    return RC_OK;
}

/* running CRC over `fit` consecutive chunks: crc(A||B) = crc(A)*x^(8|B|) + crc(B) with the
 * multiplier of a full chunk computed once per call (reference does one zlib crc32_combine per
 * chunk, src/qatzip.c:1707-1714, including its "0 restarts" rule) */
static uint32_t fold_chunk_crcs(int fmt, uint32_t crc, const uint32_t *ck, uint32_t fit, uint32_t chunk_sz, uint64_t batch_len, uint32_t xchunk)
{
    if (fmt == QZB_FMT_ZLIB) {
        /* zlib sessions carry Adler-32 per chunk; the running value is the Adler-32 of everything
         * consumed (the reference pushes Adler values through crc32_combine here, which is not a
         * checksum of anything -- DESIGN.md section 4) */
        for (uint32_t i = 0; i < fit; i++) {
            const uint64_t clen = std::min<uint64_t>(chunk_sz, batch_len - (uint64_t)i * chunk_sz);
            crc = (crc == 0) ? ck[i] : qz_adler32_combine(crc, ck[i], clen);
        }
        return crc;
    }
    for (uint32_t i = 0; i < fit; i++) {
        const uint64_t clen = std::min<uint64_t>(chunk_sz, batch_len - (uint64_t)i * chunk_sz);
        if (crc == 0) crc = ck[i];
        else crc = (clen == chunk_sz ? qz_gf2_mul(crc, xchunk) : qz_gf2_mul(crc, qz_crc_xpow8(clen))) ^ ck[i];
    }
    return crc;
}

extern "C" int qzb_engine_compress(QzbEngine *e, const QzbCompressCall *c, QzbCompressOut *o)
{
    memset(o, 0, sizeof *o);
    if (!e || !c || c->chunk_sz < 1024 || (c->chunk_sz & (c->chunk_sz - 1))) return RC_PARAMS;
    SlotGuard guard(e);
    CK(cudaSetDevice(e->device));
    uint32_t crc = c->crc_in;
    const uint32_t xchunk = c->want_crc ? qz_crc_xpow8(c->chunk_sz) : 0;
    const uint64_t per_chunk_out = (uint64_t)c->chunk_sz + (c->chunk_sz >> 7) + 256;   /* worst case incl. framing */

    if (c->src_device && c->dst_device) {
        /* device-resident: one launch per <= 1 GiB slab so piece counts stay 32-bit and scratch bounded */
        const uint64_t slab = ((uint64_t)1 << 30) / c->chunk_sz * c->chunk_sz;
        uint64_t in = 0, out = 0; int rc = RC_OK;
        Slot &s = e->slot[0];
        do {
            const uint64_t len = std::min<uint64_t>(slab, c->src_len - in);
            const int last = (in + len == c->src_len) ? c->last : 0;
            if (enqueue_compress(e, s, c, c->src + in, len, c->dst + out, c->dst_cap - out, last, &o->kernel_launches) != RC_OK) return RC_FAIL;
            CK(cudaEventSynchronize(s.ev_meta));
            float ms = 0; cudaEventElapsedTime(&ms, s.ev_k0, s.ev_k1); o->kernel_ms += ms;
            cudaEventElapsedTime(&ms, s.ev_k0, s.ev_km); o->codec_ms += ms; o->codec_launches++;
            const uint64_t *off = (const uint64_t *)s.h_meta.p;
            const uint32_t *ck = (const uint32_t *)((const uint8_t *)s.h_meta.p + align_up((size_t)(s.nchunks + 1) * 8, 16));
            uint32_t fit = 0;
            while (fit < s.nchunks && off[fit + 1] <= c->dst_cap - out) fit++;
            if (c->want_crc && c->fmt != QZB_FMT_LZ4) crc = fold_chunk_crcs(c->fmt, crc, ck, fit, c->chunk_sz, len, xchunk);
            o->nchunks += fit;
            out += off[fit];
            if (fit < s.nchunks) { in += (uint64_t)fit * c->chunk_sz; rc = RC_BUF_ERROR; break; }
            in += len;
        } while (in < c->src_len);
        o->consumed = in; o->produced = out; o->crc = crc;
        return rc;
    }
    if (c->src_device || c->dst_device) return RC_PARAMS;      /* mixed residency is not offered */

    /* host buffers: batches through two slots */
    const uint64_t batch = std::max<uint64_t>(c->chunk_sz, e->tune.batch_bytes / c->chunk_sz * c->chunk_sz);
    const uint64_t nb = c->src_len ? (c->src_len + batch - 1) / batch : 1;
    uint64_t out = 0, consumed = 0; int rc = RC_OK; bool stop = false;

    const auto host_t0 = std::chrono::steady_clock::now();
    auto host_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count(); };
    if (e->timeline) CK(cudaEventRecord(e->ev_base, e->slot[0].st));
    auto drain = [&](Slot &s) -> int {
        if (!s.busy) return RC_OK;
        CK(cudaSetDevice(s.device));
        const double host_drain0 = e->timeline ? host_ms() : 0.0;
        CK(cudaEventSynchronize(s.ev_meta));
        s.busy = false;
        const double host_meta = e->timeline ? host_ms() : 0.0;
        if (stop) { CK(cudaStreamSynchronize(s.st)); return RC_OK; }
        float ms = 0; cudaEventElapsedTime(&ms, s.ev_k0, s.ev_k1); o->kernel_ms += ms;
        cudaEventElapsedTime(&ms, s.ev_k0, s.ev_km); o->codec_ms += ms; o->codec_launches++;
        const uint64_t *off = (const uint64_t *)s.h_meta.p;
        const uint32_t *ck = (const uint32_t *)((const uint8_t *)s.h_meta.p + align_up((size_t)(s.nchunks + 1) * 8, 16));
        uint32_t fit = 0;
        while (fit < s.nchunks && off[fit + 1] <= c->dst_cap - out) fit++;
        const uint64_t bytes = off[fit];
        CK(cudaEventRecord(s.ev_d0, s.st));
        if (bytes) {
            if (c->dst_pinned) CK(cudaMemcpyAsync(c->dst + out, s.d_out.p, bytes, cudaMemcpyDeviceToHost, s.st));
            else {
                if (s.h_out.ensure(bytes) != RC_OK) return RC_FAIL;
                CK(cudaMemcpyAsync(s.h_out.p, s.d_out.p, bytes, cudaMemcpyDeviceToHost, s.st));
            }
        }
        CK(cudaEventRecord(s.ev_d1, s.st));
        if (c->want_crc && c->fmt != QZB_FMT_LZ4) crc = fold_chunk_crcs(c->fmt, crc, ck, fit, c->chunk_sz, s.in_len, xchunk);
        CK(cudaStreamSynchronize(s.st));
        { float t = 0; cudaEventElapsedTime(&t, s.ev_h0, s.ev_k0); o->h2d_ms += t; cudaEventElapsedTime(&t, s.ev_d0, s.ev_d1); o->d2h_ms += t; }
        if (e->timeline && s.device == e->device) {
            float h0 = 0, k0 = 0, km = 0, k1 = 0, d0 = 0, d1 = 0;
            cudaEventElapsedTime(&h0, e->ev_base, s.ev_h0); cudaEventElapsedTime(&k0, e->ev_base, s.ev_k0); cudaEventElapsedTime(&km, e->ev_base, s.ev_km);
            cudaEventElapsedTime(&k1, e->ev_base, s.ev_k1); cudaEventElapsedTime(&d0, e->ev_base, s.ev_d0); cudaEventElapsedTime(&d1, e->ev_base, s.ev_d1);
            fprintf(stderr, "[qzb timeline] in %5.1f MiB out %5.1f MiB | dev: h2d %.2f-%.2f codec -%.2f frame -%.2f d2h %.2f-%.2f | host: drain@%.2f meta@%.2f done@%.2f\n",
                    s.in_len / 1048576.0, bytes / 1048576.0, h0, k0, km, k1, d0, d1, host_drain0, host_meta, host_ms());
        }
        if (bytes && !c->dst_pinned) memcpy(c->dst + out, s.h_out.p, bytes);
        out += bytes; o->nchunks += fit;
        if (fit < s.nchunks) { consumed += (uint64_t)fit * c->chunk_sz; rc = RC_BUF_ERROR; stop = true; }
        else consumed += s.in_len;
        return RC_OK;
    };

    /* Ring of NS slots, NSLOT on each of the engine's devices, dealt device by device: batch b goes to device b % G.  It is
     * issued as soon as that slot's previous tenant (batch b - NS) has been drained; after issuing b the oldest undrained
     * batch is drained, so that while the host waits for it younger batches are already queued on every device.  Draining
     * in batch order is what lays the output end to end and folds the checksums in order, whichever GPU made them. */
    const int G = (int)e->devices.size();
    const int NS = QzbEngine::NSLOT * G;
    auto slot_of = [&](uint64_t b) -> Slot & { return e->slot[(b % G) * QzbEngine::NSLOT + ((b / G) % QzbEngine::NSLOT)]; };
    uint64_t next_drain = 0;
    /* batch sizes ramp up 16, 32, 64 ... MiB (QZB200_FIRST_MB, QZB200_BATCH_MB): the first kernel starts after a short copy, and a call's
     * unavoidable fill/drain tail (calls are synchronous) stays small against its steady state */
    uint64_t in_off_next = 0, issued = 0;
    for (uint64_t b = 0; (in_off_next < c->src_len || b == 0) && !stop; b++) {
        while (next_drain + NS <= b) { if (drain(slot_of(next_drain)) != RC_OK) return RC_FAIL; next_drain++; }
        if (stop) break;
        Slot &s = slot_of(b);
        CK(cudaSetDevice(s.device));
        const uint64_t first = std::max<uint64_t>(c->chunk_sz, e->tune.first_batch_bytes / c->chunk_sz * c->chunk_sz);
        const uint64_t ramp = first << std::min<uint64_t>(b / G, 10);
        const uint64_t in_off = in_off_next, left = c->src_len - in_off;
        uint64_t len = std::min<uint64_t>(std::min(batch, ramp), left);
        /* optional taper (QZB200_TAPER=1): no batch takes more than half of what is left, so the last
         * kernel and the last copy back are short.  Off by default: on B200 the extra batches cost more
         * than the shorter tail saves (profiles/r01_e2e_batching.md) */
        if (e->tune.taper && left > first) len = std::min<uint64_t>(len, std::max<uint64_t>(first, left / 2 / c->chunk_sz * c->chunk_sz));
        in_off_next = in_off + len;
        const uint32_t nch = len ? (uint32_t)((len + c->chunk_sz - 1) / c->chunk_sz) : 1u;
        if (s.d_in.ensure(len + 64) != RC_OK || s.d_out.ensure((uint64_t)nch * per_chunk_out) != RC_OK) return RC_FAIL;
        CK(cudaEventRecord(s.ev_h0, s.st));
        if (len) {
            if (c->src_pinned) CK(cudaMemcpyAsync(s.d_in.p, c->src + in_off, len, cudaMemcpyHostToDevice, s.st));
            else {
                if (s.h_in.ensure(len) != RC_OK) return RC_FAIL;
                memcpy(s.h_in.p, c->src + in_off, len);
                CK(cudaMemcpyAsync(s.d_in.p, s.h_in.p, len, cudaMemcpyHostToDevice, s.st));
            }
        }
        const int last = (in_off + len == c->src_len) ? c->last : 0;
        s.busy = true; s.in_off = in_off; s.in_len = len;
        if (enqueue_compress(e, s, c, (const uint8_t *)s.d_in.p, len, (uint8_t *)s.d_out.p, s.d_out.cap, last, &o->kernel_launches) != RC_OK) return RC_FAIL;
        issued = b + 1;
        if (issued - next_drain >= (uint64_t)NS) {       /* ring full: retire the oldest while NS-1 younger ones are queued */
            if (drain(slot_of(next_drain)) != RC_OK) return RC_FAIL; next_drain++;
        }
    }
    for (; next_drain < issued; next_drain++) if (drain(slot_of(next_drain)) != RC_OK) return RC_FAIL;
    o->consumed = consumed; o->produced = out; o->crc = crc;
    return rc;
}

/* ------------------------------------------------------------------ decompress */
static inline uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }

struct ParsedMember {
    QzbMember m;
    uint64_t unit_start;       /* offset of the member/frame header in the call's src */
    uint32_t hdr_len, ftr_len;
    bool sized;                /* payload length and output size known up front */
    bool speculative;          /* extent guessed by the gzip magic scan: verified by the decode */
    bool force_seq;            /* guessed extent is implausible: decode alone, device finds the end */
};

/* RFC 1952 header walk.  Returns header length, 0 if more bytes are needed, -1 if not gzip. */
static long gzip_header_len(const uint8_t *p, uint64_t avail, bool *has_qz, uint32_t *qz_src, uint32_t *qz_dst)
{
    *has_qz = false;
    if (avail < 10) return 0;
    if (p[0] != 0x1f || p[1] != 0x8b || p[2] != 8) return -1;
    const uint8_t flg = p[3]; uint64_t h = 10;
    if (flg & 0xe0) return -1;
    if (flg & 4) {
        if (avail < h + 2) return 0;
        const uint32_t xlen = p[h] | (uint32_t)p[h + 1] << 8;
        if (avail < h + 2 + xlen) return 0;
        /* the fixed-shape 'QZ' sub-field: reference src/qatzip_gzip.c:200-226 */
        if (xlen == 12 && p[h + 2] == 'Q' && p[h + 3] == 'Z' && p[h + 4] == 8 && p[h + 5] == 0) { *has_qz = true; *qz_src = rd32(p + h + 6); *qz_dst = rd32(p + h + 10); }
        h += 2 + xlen;
    }
    if (flg & 8) { while (h < avail && p[h]) h++; if (h >= avail) return 0; h++; }
    if (flg & 16) { while (h < avail && p[h]) h++; if (h >= avail) return 0; h++; }
    if (flg & 2) h += 2;
    if (h > avail) return 0;
    return (long)h;
}

/* Positions in p[lo, hi) that look like the start of a gzip member: magic, method, reserved flag bits clear,
 * XFL and OS values that exist (RFC 1952 2.3.1).  The reference finds member ends with one linear scan
 * per member (src/qatzip_gzip.c:244-261); here the whole input is scanned once, sliced over host threads,
 * and the member walk looks boundaries up in the sorted result. */
static bool gzip_member_start(const uint8_t *p, uint64_t q, uint64_t n)
{
    return q + 10 <= n && p[q] == 0x1f && p[q + 1] == 0x8b && p[q + 2] == 8 && (p[q + 3] & 0xe0) == 0 &&
           (p[q + 8] == 0 || p[q + 8] == 2 || p[q + 8] == 4) && (p[q + 9] <= 13 || p[q + 9] == 255);
}
static void gzip_scan_candidates(const uint8_t *p, uint64_t lo, uint64_t n, std::vector<uint64_t> &out)
{
    out.clear();
    if (n < lo + 10) return;
    const uint64_t span = n - lo;
    /* one slice of at least 8 MiB per thread, at most 64 threads (a 1.5 GiB call: 3 ms instead of 18 with sixteen) */
    unsigned T = std::thread::hardware_concurrency(); if (T == 0) T = 4; if (T > 64) T = 64;
    if (span / (8u << 20) < T) T = (unsigned)std::max<uint64_t>(1, span / (8u << 20));
    std::vector<std::vector<uint64_t>> part(T);
    auto work = [&](unsigned t) {
        const uint64_t a = lo + span * t / T, b = lo + span * (t + 1) / T;
        uint64_t q = a;
#if defined(__SSE2__)
        /* 16 positions a step: 1f at q and 8b at q + 1 (one in 65 536 positions of compressed data passes) */
        const __m128i c1f = _mm_set1_epi8(0x1f), c8b = _mm_set1_epi8((char)0x8b);
        while (q + 17 <= b && q + 17 <= n) {
            const __m128i x = _mm_loadu_si128((const __m128i *)(p + q)), y = _mm_loadu_si128((const __m128i *)(p + q + 1));
            unsigned m = (unsigned)_mm_movemask_epi8(_mm_and_si128(_mm_cmpeq_epi8(x, c1f), _mm_cmpeq_epi8(y, c8b)));
            while (m) { const unsigned k = (unsigned)__builtin_ctz(m); m &= m - 1; if (gzip_member_start(p, q + k, n)) part[t].push_back(q + k); }
            q += 16;
        }
#endif
        while (q < b) {
            const uint8_t *f = (const uint8_t *)memchr(p + q, 0x1f, b - q);
            if (!f) break;
            q = (uint64_t)(f - p);
            if (gzip_member_start(p, q, n)) part[t].push_back(q);
            q++;
        }
    };
    if (T == 1) work(0);
    else { std::vector<std::thread> th; for (unsigned t = 0; t < T; t++) th.emplace_back(work, t); for (auto &x : th) x.join(); }
    for (auto &v : part) out.insert(out.end(), v.begin(), v.end());
}

extern "C" int qzb_engine_decompress(QzbEngine *e, const QzbDecompressCall *c, QzbDecompressOut *o)
{
    std::vector<uint64_t> gz_cand; bool gz_scanned = false;       /* absolute offsets of plausible gzip member starts */
    const auto host_t0 = std::chrono::steady_clock::now();
    auto host_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count(); };
    memset(o, 0, sizeof *o);
    if (!e || !c) return RC_PARAMS;
    CK(cudaSetDevice(e->device));
    const uint8_t *hsrc = c->src_device ? c->src_host_view : c->src;
    if (!hsrc) return RC_PARAMS;
    if (c->src_device != c->dst_device) return RC_PARAMS;
    SlotGuard guard(e);
    const bool lz4 = (c->fmt == QZB_FMT_LZ4);
    uint64_t cur_in = 0, cur_out = 0;      /* everything before these offsets is decoded and delivered */
    uint64_t zlib_window = e->tune.zlib_window_bytes;       /* span searched for zlib stream starts per round */

    /* one kernel launch over units[first, first+count); outputs either at their final offsets
     * (relative to out_base) or, when staged, in private regions of the slot's d_out */
    std::vector<ParsedMember> units;
    bool size_only = false;                /* zlib member discovery: lengths only, nothing written */
    auto run = [&](Slot &s, size_t first, size_t count, uint64_t span_src, uint64_t span_len, uint64_t out_base, uint64_t out_len,
                   bool staged_out) -> int {
        const uint8_t *d_src; uint8_t *d_dst = nullptr;
        if (c->src_device) { d_src = c->src + span_src; if (!staged_out) d_dst = c->dst + out_base; }
        else {
            if (s.d_in.ensure(span_len + 64) != RC_OK) return RC_FAIL;
            if (c->src_pinned) CK(cudaMemcpyAsync(s.d_in.p, c->src + span_src, span_len, cudaMemcpyHostToDevice, s.st));
            else { if (s.h_in.ensure(span_len) != RC_OK) return RC_FAIL; memcpy(s.h_in.p, c->src + span_src, span_len);
                   CK(cudaMemcpyAsync(s.d_in.p, s.h_in.p, span_len, cudaMemcpyHostToDevice, s.st)); }
            d_src = (const uint8_t *)s.d_in.p;
        }
        if (!d_dst) { if (s.d_out.ensure(out_len + 64) != RC_OK) return RC_FAIL; d_dst = (uint8_t *)s.d_out.p; }
        /* the member table, and behind it the order the members are handed out in (largest payload first) */
        const size_t order_off = align_up(count * sizeof(QzbMember), 16);
        if (s.h_members.ensure(order_off + count * 4) != RC_OK || s.d_members.ensure(order_off + count * 4) != RC_OK) return RC_FAIL;
        if (s.h_results.ensure(count * sizeof(QzbMemberResult)) != RC_OK || s.d_results.ensure(count * sizeof(QzbMemberResult) + 16) != RC_OK) return RC_FAIL;
        QzbMember *hm = (QzbMember *)s.h_members.p;
        uint64_t stage_off = 0;
        for (size_t i = 0; i < count; i++) {
            hm[i] = units[first + i].m;
            hm[i].src_off -= span_src;
            if (size_only) hm[i].dst_off = 0;
            else if (staged_out) { hm[i].dst_off = stage_off; stage_off += align_up(hm[i].dst_cap, 16); }
            else hm[i].dst_off -= out_base;
        }
        bool ordered = false;
        if (!lz4 && count > 64) {
            uint32_t lo = 0xffffffffu, hi = 0;
            for (size_t i = 0; i < count; i++) { lo = std::min(lo, hm[i].src_len); hi = std::max(hi, hm[i].src_len); }
            if (hi > 2 * (uint64_t)lo) {
                uint32_t *ord = (uint32_t *)((uint8_t *)s.h_members.p + order_off);
                /* longest payload first, equal lengths in member order: one sort over (inverted length, index) keys */
                std::vector<uint64_t> keys(count);
                for (size_t i = 0; i < count; i++) keys[i] = (uint64_t)(0xffffffffu - hm[i].src_len) << 32 | (uint32_t)i;
                std::sort(keys.begin(), keys.end());
                for (size_t i = 0; i < count; i++) ord[i] = (uint32_t)keys[i];
                ordered = true;
            }
        }
        CK(cudaMemcpyAsync(s.d_members.p, hm, ordered ? order_off + count * 4 : count * sizeof(QzbMember), cudaMemcpyHostToDevice, s.st));
        uint32_t *ticket = (uint32_t *)((uint8_t *)s.d_results.p + count * sizeof(QzbMemberResult));
        CK(cudaMemsetAsync(ticket, 0, 16, s.st));
        QzbDecompressJob job; memset(&job, 0, sizeof job);
        job.src = d_src; job.dst = d_dst; job.members = (const QzbMember *)s.d_members.p; job.results = (QzbMemberResult *)s.d_results.p;
        job.nmembers = (uint32_t)count; job.fmt = c->fmt; job.ticket = ticket; job.size_only = size_only ? 1 : 0;
        if (ordered) job.order = (const uint32_t *)((const uint8_t *)s.d_members.p + order_off);
        const int dpw = e->tune.inflate_dpw;
        const size_t slots_per_cta = lz4 ? 8 : (size_t)qzb_inflate_cta_slots(dpw);
        const int ctas_per_sm = lz4 ? 8 : (int)std::max<size_t>(1, std::min<size_t>(2048 / qzb_inflate_cta_threads(dpw), (228 * 1024) / (qzb_inflate_smem_bytes(dpw) + 1024 + 1200)));
        const int grid = (int)std::min<size_t>((count + slots_per_cta - 1) / slots_per_cta, (size_t)e->sm_count * ctas_per_sm);
        CK(cudaEventRecord(s.ev_k0, s.st));
        if (lz4) CK(qzb_launch_lz4_decompress(&job, grid, s.st));
        else CK(qzb_launch_inflate(&job, dpw, grid, s.st));
        CK(cudaEventRecord(s.ev_k1, s.st));
        o->kernel_launches += 1;
        CK(cudaMemcpyAsync(s.h_results.p, s.d_results.p, count * sizeof(QzbMemberResult), cudaMemcpyDeviceToHost, s.st));
        CK(cudaEventRecord(s.ev_meta, s.st));
        s.busy = true; s.first_member = first; s.nmembers = count; s.span_src = span_src; s.span_len = span_len; s.out_base = out_base; s.out_len = out_len;
        return RC_OK;
    };
    auto status_rc = [&](const QzbMemberResult &r, bool sized) -> int {
        if (r.status == QZB_ST_OK) return RC_OK;
        if (r.status == QZB_ST_OUT_FULL) return sized ? RC_DATA_ERROR : RC_BUF_ERROR;
        return RC_DATA_ERROR;
    };
    /* decode ONE unit whose extent is not trusted / not known: payload = everything that is left,
     * output into staging; the device reports how much it really used and made */
    auto run_single = [&](ParsedMember u, uint32_t src_len, uint32_t cap, QzbMemberResult *res, const uint8_t **staged) -> int {
        Slot &s = e->slot[0];
        units.assign(1, u);
        units[0].m.src_len = src_len; units[0].m.exact_len &= ~1u; units[0].m.dst_cap = cap; units[0].m.exact_out = 0; units[0].m.check_cksum = 0;
        if (run(s, 0, 1, u.m.src_off, src_len, 0, align_up(cap, 16), true) != RC_OK) return RC_FAIL;
        s.busy = false;
        CK(cudaEventSynchronize(s.ev_meta));
        float ms = 0; cudaEventElapsedTime(&ms, s.ev_k0, s.ev_k1); o->kernel_ms += ms;
        *res = *(const QzbMemberResult *)s.h_results.p;
        *staged = (const uint8_t *)s.d_out.p;
        return RC_OK;
    };
    auto deliver = [&](const uint8_t *staged, uint64_t nbytes) -> int {
        if (!nbytes) return RC_OK;
        Slot &s = e->slot[0];
        CK(cudaMemcpyAsync(c->dst + cur_out, staged, nbytes, c->dst_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s.st));
        CK(cudaStreamSynchronize(s.st));
        return RC_OK;
    };

    int final_rc = RC_OK;
    while (cur_in < c->src_len) {
        /* ---- pass 1 (host): walk unit headers from cur_in, like checkHeader does per request ---- */
        units.clear();
        int parse_rc = RC_OK;
        uint64_t in = cur_in, out = cur_out;
        if (c->fmt == QZB_FMT_ZLIB) {
            /* A zlib stream carries neither its compressed nor its uncompressed size (the reference
             * submits "everything that is left" and reads `consumed` back, src/qatzip_utils.c:1309-1315,
             * one stream at a time).  Here every offset that passes the header test
             * (src/qatzip_gzip.c:283-306) is decoded for LENGTHS ONLY in one launch; walking the
             * (consumed, produced) links from the first stream yields the real members, which then go
             * through the ordinary sized path with exact lengths and their Adler-32. */
            uint64_t win = std::min<uint64_t>(c->src_len - cur_in, c->src_device ? 0xfffffff0ull : zlib_window);
            const uint8_t *p = hsrc + cur_in;
            auto hdr_ok = [&](uint64_t q) { return (p[q] & 0x0f) == 8 && (p[q] >> 4) <= 7 && !(p[q + 1] & 0x20) && ((uint32_t)p[q] * 256u + p[q + 1]) % 31u == 0; };
            std::vector<uint64_t> cand;
            if (win >= 2 && hdr_ok(0)) {
                cand.push_back(0);
                if (!c->stop_at_first && win >= 7) {
                    /* every offset is tested (one in ~2000 passes): sliced over host threads like the gzip member scan */
                    const uint64_t lo = 1, hi = win - 5;
                    unsigned T = std::thread::hardware_concurrency(); if (T == 0) T = 4; if (T > 64) T = 64;
                    if ((hi - lo) / (4u << 20) < T) T = (unsigned)std::max<uint64_t>(1, (hi - lo) / (4u << 20));
                    std::vector<std::vector<uint64_t>> part(T);
                    auto work = [&](unsigned t) {
                        const uint64_t a = lo + (hi - lo) * t / T, b = lo + (hi - lo) * (t + 1) / T;
                        for (uint64_t q = a; q < b; q++) if (hdr_ok(q)) part[t].push_back(q);
                    };
                    if (T == 1) work(0);
                    else { std::vector<std::thread> th; for (unsigned t = 0; t < T; t++) th.emplace_back(work, t); for (auto &x : th) x.join(); }
                    for (auto &v : part) cand.insert(cand.end(), v.begin(), v.end());
                }
            }
            if (cand.empty()) { final_rc = (win < 2) ? RC_DATA_ERROR : RC_FAIL; break; }
            for (uint64_t q : cand) {
                ParsedMember u; memset(&u, 0, sizeof u);
                u.unit_start = cur_in + q; u.m.src_off = cur_in + q + 2; u.m.src_len = (uint32_t)(win - q - 2); u.m.dst_cap = (uint32_t)std::min<uint64_t>(c->dst_cap - cur_out, 0xfffffff0ull);   /* more than fits cannot be delivered anyway */
                units.push_back(u);
            }
            Slot &s0 = e->slot[0];
            size_only = true;
            const int rr = run(s0, 0, units.size(), cur_in, win, 0, 0, true);
            size_only = false;
            if (rr != RC_OK) return RC_FAIL;
            s0.busy = false;
            CK(cudaEventSynchronize(s0.ev_meta));
            { float ms = 0; cudaEventElapsedTime(&ms, s0.ev_k0, s0.ev_k1); o->kernel_ms += ms; }
            const QzbMemberResult *res = (const QzbMemberResult *)s0.h_results.p;
            std::vector<ParsedMember> chain;
            uint64_t pos = 0; bool need_more = false;
            while (pos < win) {
                auto it = std::lower_bound(cand.begin(), cand.end(), pos);
                if (it == cand.end() || *it != pos) {
                    /* too close to the window end to have been tried, or not a zlib header at all */
                    const bool more_input = cur_in + win < c->src_len;
                    if (win - pos < 2) { if (more_input) need_more = true; else parse_rc = RC_DATA_ERROR; }
                    else if (!hdr_ok(pos)) parse_rc = RC_FAIL;
                    else if (more_input) need_more = true;
                    else parse_rc = RC_DATA_ERROR;
                    break;
                }
                const QzbMemberResult &r = res[it - cand.begin()];
                /* ran off the end of the window (the reader feeds zero bits there, so this can also surface as a data error) */
                const uint64_t avail = win - pos - 2;
                const bool cut = (r.status == QZB_ST_IN_TRUNC) || (r.status != QZB_ST_OK && r.status != QZB_ST_OUT_FULL && (uint64_t)r.consumed + 4 >= avail) ||
                                 (r.status == QZB_ST_OK && (!r.saw_final || (uint64_t)r.consumed + 4 > avail));
                if (cut) { if (cur_in + win < c->src_len) need_more = true; else parse_rc = RC_DATA_ERROR; break; }
                if (r.status != QZB_ST_OK) { parse_rc = RC_DATA_ERROR; break; }
                if ((uint64_t)r.produced > c->dst_cap - out) { parse_rc = RC_BUF_ERROR; break; }
                const uint8_t *f = p + pos + 2 + r.consumed;
                ParsedMember u; memset(&u, 0, sizeof u);
                u.unit_start = cur_in + pos; u.hdr_len = 2; u.ftr_len = 4;
                u.m.src_off = cur_in + pos + 2; u.m.src_len = r.consumed; u.m.exact_len = 1;
                u.m.dst_off = out; u.m.dst_cap = r.produced; u.m.exact_out = 1;
                u.m.expect_cksum = (uint32_t)f[0] << 24 | (uint32_t)f[1] << 16 | (uint32_t)f[2] << 8 | f[3]; u.m.check_cksum = 1; u.sized = true;
                chain.push_back(u);
                out += r.produced; pos += 2 + (uint64_t)r.consumed + 4;
                if (c->stop_at_first) break;
            }
            if (getenv("QZB200_DEBUG")) {
                const QzbMemberResult &r0 = res[0];
                fprintf(stderr, "[qatzip_b200] zlib discovery at %llu: window %llu, %zu candidates, chain %zu, stopped at +%llu (need_more %d, rc %d); first: status %u consumed %u produced %u final %u\n",
                        (unsigned long long)cur_in, (unsigned long long)win, cand.size(), chain.size(), (unsigned long long)pos, (int)need_more, parse_rc,
                        r0.status, r0.consumed, r0.produced, r0.saw_final);
                if (pos < win) { auto it = std::lower_bound(cand.begin(), cand.end(), pos); if (it != cand.end() && *it == pos) { const QzbMemberResult &r = res[it - cand.begin()];
                    fprintf(stderr, "[qatzip_b200]   stopping candidate: status %u consumed %u produced %u final %u avail %llu\n", r.status, r.consumed, r.produced, r.saw_final, (unsigned long long)(win - pos - 2)); } }
            }
            if (chain.empty() && need_more) {
                /* not even one whole stream inside the window: widen it */
                if (c->src_device || zlib_window >= c->src_len - cur_in) { final_rc = RC_DATA_ERROR; break; }
                zlib_window *= 4;
                continue;
            }
            units.swap(chain);
            in = c->src_len;                   /* skip the header walk below */
        }
        while (in < c->src_len) {
            const uint8_t *p = hsrc + in; const uint64_t avail = c->src_len - in;
            ParsedMember u; memset(&u, 0, sizeof u); u.unit_start = in;
            if (c->fmt == QZB_FMT_GZIP || c->fmt == QZB_FMT_GZIP_EXT) {
                bool has_qz; uint32_t qsrc = 0, qdst = 0;
                long h = gzip_header_len(p, avail, &has_qz, &qsrc, &qdst);
                if (h < 0) { parse_rc = RC_FAIL; break; }
                if (h == 0) { parse_rc = RC_DATA_ERROR; break; }
                u.hdr_len = (uint32_t)h; u.ftr_len = 8;
                uint64_t payload;
                if (has_qz) payload = qdst;
                else {
                    /* next member by magic scan, footer right before it: reference src/qatzip_gzip.c:244-261.
                     * The guess is verified by the decode (exact length, ISIZE, CRC) and replaced by a
                     * sequential decode of this member if it does not hold. */
                    if (!gz_scanned) {
                        const double t0 = e->timeline ? host_ms() : 0.0;
                        bool on_device = false;
                        if (c->src_device && c->src_len - in < 0xffffffffull) {
                            /* the compressed bytes are in device memory: scan them there (a gigabyte in well under a
                             * millisecond; the host needs ~10 ms on sixteen threads), bring the offsets back and sort them */
                            Slot &s0 = e->slot[0];
                            const uint64_t span = c->src_len - in;
                            const uint32_t lcap = (uint32_t)std::min<uint64_t>(span / 256 + 65536, 0x3fffffffu);
                            if (s0.d_meta.ensure(16 + (size_t)lcap * 4) == RC_OK && s0.h_meta.ensure(16) == RC_OK) {
                                uint32_t *d_count = (uint32_t *)s0.d_meta.p, *d_list = d_count + 4;
                                CK(cudaMemsetAsync(d_count, 0, 16, s0.st));
                                CK(qzb_launch_gzip_scan(c->src, in, c->src_len, d_list, lcap, d_count, e->sm_count * 8, s0.st));
                                CK(cudaMemcpyAsync(s0.h_meta.p, d_count, 4, cudaMemcpyDeviceToHost, s0.st));
                                CK(cudaStreamSynchronize(s0.st));
                                const uint32_t found = *(const uint32_t *)s0.h_meta.p;
                                if (found <= lcap && s0.h_meta.ensure(16 + (size_t)found * 4) == RC_OK) {
                                    if (found) { CK(cudaMemcpyAsync((uint8_t *)s0.h_meta.p + 16, d_list, (size_t)found * 4, cudaMemcpyDeviceToHost, s0.st)); CK(cudaStreamSynchronize(s0.st)); }
                                    uint32_t *hl = (uint32_t *)((uint8_t *)s0.h_meta.p + 16);
                                    std::sort(hl, hl + found);
                                    gz_cand.resize(found);
                                    for (uint32_t i = 0; i < found; i++) gz_cand[i] = in + hl[i];
                                    on_device = true;
                                }
                            }
                        }
                        if (!on_device) gzip_scan_candidates(hsrc, in, c->src_len, gz_cand);
                        gz_scanned = true;
                        if (e->timeline) fprintf(stderr, "[qzb timeline] inflate: member scan of %.1f MiB (%s): %.2f ms, %zu candidates\n", (double)(c->src_len - in) / 1048576.0,
                                                 on_device ? "device" : "host", host_ms() - t0, gz_cand.size());
                    }
                    /* first candidate that leaves room for this member's footer and has a believable ISIZE in front */
                    uint64_t q = 0; bool found = false;
                    auto it0 = std::lower_bound(gz_cand.begin(), gz_cand.end(), in + (uint64_t)h + 8);
                    /* the walk reads one header and one footer per member, each a cache miss on the host view: ask for the
                     * ones a few members ahead now */
                    if (gz_cand.end() - it0 > 6) { const uint8_t *pf = hsrc + *(it0 + 6); __builtin_prefetch(pf - 8); __builtin_prefetch(pf + 24); }
                    for (auto it = it0; it != gz_cand.end(); ++it) {
                        q = *it - in;
                        if ((uint64_t)rd32(p + q - 4) <= (q - h - 8) * 1032 + 64) { found = true; break; }
                    }
                    const uint64_t end = found ? q : avail;
                    if (end < (uint64_t)h + 8) { parse_rc = RC_DATA_ERROR; break; }
                    payload = end - h - 8;
                    u.speculative = true;
                }
                if ((uint64_t)h + payload + 8 > avail) { parse_rc = RC_DATA_ERROR; break; }
                if (payload > 0xfffffff0ull) { parse_rc = RC_FAIL; break; }
                const uint8_t *ftr = p + h + payload;
                const uint32_t isize = has_qz ? qsrc : rd32(ftr + 4);
                if ((uint64_t)isize > c->dst_cap - out) {
                    if (!u.speculative) { parse_rc = RC_BUF_ERROR; break; }
                    /* the size came from a guessed footer position: do not trust it, let the device find the end */
                    u.m.src_off = in + h; u.m.src_len = (uint32_t)payload; u.m.exact_len = 1; u.m.dst_off = out; u.m.dst_cap = 0;
                    u.sized = true; u.force_seq = true;
                    units.push_back(u);
                    break;
                }
                u.m.src_off = in + h; u.m.src_len = (uint32_t)payload; u.m.exact_len = 1;
                u.m.dst_off = out; u.m.dst_cap = isize; u.m.exact_out = 1;
                u.m.expect_cksum = rd32(ftr); u.m.check_cksum = 1; u.sized = true;
                in += h + payload + 8; out += isize;
            } else if (c->fmt == QZB_FMT_4B) {
                if (avail < 4) { parse_rc = RC_DATA_ERROR; break; }
                const uint32_t blk = rd32(p);
                if (4 + (uint64_t)blk > avail) { parse_rc = RC_DATA_ERROR; break; }
                u.hdr_len = 4; u.m.src_off = in + 4; u.m.src_len = blk; u.m.exact_len = 1;
                u.m.dst_cap = c->chunk_sz; u.m.exact_out = 0; u.sized = false;
                in += 4 + (uint64_t)blk;
            } else if (c->fmt == QZB_FMT_RAW) {
                if (avail > 0xfffffff0ull) { parse_rc = RC_FAIL; break; }
                u.m.src_off = in; u.m.src_len = (uint32_t)avail; u.m.exact_len = 0;
                /* deflate cannot expand more than 1032:1, which bounds the staging a raw stream needs */
                u.m.dst_cap = (uint32_t)std::min<uint64_t>(std::min<uint64_t>(c->dst_cap - out, avail * 1032 + 64), 0xfffffff0ull); u.sized = false;
                in += avail;
            } else if (lz4) {
                /* frame header: reference src/qatzip_lz4.c:62-102 (verify), :145-173 (walk blocks to EndMark) */
                if (avail < 7) { parse_rc = RC_DATA_ERROR; break; }
                if (rd32(p) != 0x184D2204u) { parse_rc = RC_FAIL; break; }
                const uint8_t flg = p[4];
                if ((flg >> 6) != 1) { parse_rc = RC_FAIL; break; }
                uint64_t h = 6 + ((flg & 8) ? 8 : 0) + ((flg & 1) ? 4 : 0) + 1;
                if (avail < h + 4) { parse_rc = RC_DATA_ERROR; break; }
                if (p[h - 1] != (uint8_t)(qz_xxh32(p + 4, (size_t)(h - 5), 0) >> 8)) { parse_rc = RC_DATA_ERROR; break; }
                uint64_t q = h; bool ok = false;
                while (q + 4 <= avail) {
                    const uint32_t bh = rd32(p + q);
                    if (bh == 0) { ok = true; break; }
                    q += 4 + (uint64_t)(bh & 0x7fffffffu) + ((flg & 0x10) ? 4 : 0);
                }
                if (!ok) { parse_rc = RC_DATA_ERROR; break; }
                const uint32_t ftr = 4 + ((flg & 4) ? 4 : 0);
                if (q + ftr > avail) { parse_rc = RC_DATA_ERROR; break; }
                if (q - h > 0xfffffff0ull) { parse_rc = RC_FAIL; break; }
                u.hdr_len = (uint32_t)h; u.ftr_len = ftr;
                u.m.src_off = in + h; u.m.src_len = (uint32_t)(q - h);
                /* bit 0: payload length is exact; bit 1: per-block checksums present (kernel skips them) */
                u.m.exact_len = 1u | ((flg & 0x10) ? 2u : 0u);
                u.m.check_cksum = (flg & 4) ? 1 : 0; u.m.expect_cksum = (flg & 4) ? rd32(p + q + 4) : 0;
                if (flg & 8) {
                    const uint64_t cs = (uint64_t)rd32(p + 6) | (uint64_t)rd32(p + 10) << 32;
                    if (cs > c->dst_cap - out) { parse_rc = RC_BUF_ERROR; break; }
                    if (cs > 0xfffffff0ull) { parse_rc = RC_FAIL; break; }
                    u.m.dst_off = out; u.m.dst_cap = (uint32_t)cs; u.m.exact_out = 1; u.sized = true; out += cs;
                } else { u.m.dst_cap = (uint32_t)std::min<uint64_t>(std::min<uint64_t>(c->dst_cap - out, (q - h) * 255 + 64), 0xfffffff0ull); u.sized = false; }
                in += q + ftr;
            } else return RC_PARAMS;
            units.push_back(u);
            if (c->stop_at_first) break;
            if (!u.sized && c->fmt != QZB_FMT_4B) break;     /* where the next unit's output starts is not known yet */
        }
        if (units.empty()) { final_rc = parse_rc; break; }

        if (e->timeline) fprintf(stderr, "[qzb timeline] inflate: %zu units parsed @%.2f ms\n", units.size(), host_ms());
        /* ---- pass 2 (device) ---- */
        const bool all_sized = std::all_of(units.begin(), units.end(), [](const ParsedMember &u) { return u.sized; });
        int rc2 = RC_OK; long failed_unit = -1;
        if (all_sized) {
            /* members know where they go: batch them, two slots in flight */
            /* one warp per member: a launch needs ~7000 members in flight to fill 148 SMs, so batches are
             * cut by member count first and by bytes second */
            /* host buffers: batches small enough to pipeline copies; device-resident: one launch as wide as possible */
            const uint64_t bin = c->src_device ? ((uint64_t)1 << 40) : e->tune.inflate_batch_bytes, bout = bin * 4;      /* device-resident: one launch per call */
            size_t i = 0, issued = 0; bool stop = false;
            long seq_unit = -1;
            size_t nsized = units.size();
            if (units.back().force_seq) { seq_unit = (long)units.size() - 1; nsized--; }
            auto drain = [&](Slot &s) -> int {
                if (!s.busy) return RC_OK;
                s.busy = false;
                const double t_drain = e->timeline ? host_ms() : 0.0;
                CK(cudaEventSynchronize(s.ev_meta));
                if (stop) { CK(cudaStreamSynchronize(s.st)); return RC_OK; }
                float ms = 0; cudaEventElapsedTime(&ms, s.ev_k0, s.ev_k1); o->kernel_ms += ms;
                const double t_meta = e->timeline ? host_ms() : 0.0;
                auto tl_print = [&]() { if (e->timeline) fprintf(stderr, "[qzb timeline] inflate batch: %zu members, out %.1f MiB | kernel %.2f ms | host: drain@%.2f results@%.2f delivered@%.2f\n",
                                                                s.nmembers, (double)s.out_len / 1048576.0, ms, t_drain, t_meta, host_ms()); };
                const QzbMemberResult *r = (const QzbMemberResult *)s.h_results.p;
                size_t good = 0;
                while (good < s.nmembers && r[good].status == QZB_ST_OK) good++;
                uint64_t obytes = 0;
                if (good) { const ParsedMember &lu = units[s.first_member + good - 1]; obytes = lu.m.dst_off + lu.m.dst_cap - s.out_base; }
                if (!c->dst_device && obytes) {
                    if (c->dst_pinned) CK(cudaMemcpyAsync(c->dst + s.out_base, s.d_out.p, obytes, cudaMemcpyDeviceToHost, s.st));
                    else { if (s.h_out.ensure(obytes) != RC_OK) return RC_FAIL; CK(cudaMemcpyAsync(s.h_out.p, s.d_out.p, obytes, cudaMemcpyDeviceToHost, s.st)); }
                    CK(cudaStreamSynchronize(s.st));
                    if (!c->dst_pinned) memcpy(c->dst + s.out_base, s.h_out.p, obytes);
                }
                if (good) { const ParsedMember &lu = units[s.first_member + good - 1]; cur_in = lu.unit_start + lu.hdr_len + lu.m.src_len + lu.ftr_len; cur_out = s.out_base + obytes; o->nmembers += (uint32_t)good; }
                if (good < s.nmembers) {
                    rc2 = status_rc(r[good], true); failed_unit = (long)(s.first_member + good); stop = true;
                    if (getenv("QZB200_DEBUG")) fprintf(stderr, "[qatzip_b200] member %ld failed: status %u consumed %u/%u produced %u/%u cksum %08x/%08x\n", failed_unit, r[good].status,
                                                        r[good].consumed, units[(size_t)failed_unit].m.src_len, r[good].produced, units[(size_t)failed_unit].m.dst_cap, r[good].cksum, units[(size_t)failed_unit].m.expect_cksum);
                }
                tl_print();
                return RC_OK;
            };
            constexpr int NS = QzbEngine::NSLOT;
            size_t next_drain = 0;
            while (i < nsized && !stop) {
                size_t j = i; uint64_t sin = 0, sout = 0;
                while (j < nsized) {
                    const uint64_t ulen = units[j].hdr_len + (uint64_t)units[j].m.src_len + units[j].ftr_len;
                    if (j > i && (sin + ulen > bin || sout + units[j].m.dst_cap > bout)) break;
                    sin += ulen; sout += units[j].m.dst_cap; j++;
                }
                while (next_drain + NS <= issued) { if (drain(e->slot[next_drain % NS]) != RC_OK) return RC_FAIL; next_drain++; }
                if (stop) break;
                Slot &s = e->slot[issued % NS];
                if (run(s, i, j - i, units[i].unit_start, sin, units[i].m.dst_off, sout, false) != RC_OK) return RC_FAIL;
                if (e->timeline) fprintf(stderr, "[qzb timeline] inflate: batch of %zu issued @%.2f ms\n", j - i, host_ms());
                issued++; i = j;
                if (issued - next_drain >= (size_t)NS) { if (drain(e->slot[next_drain % NS]) != RC_OK) return RC_FAIL; next_drain++; }
            }
            for (int k = 0; k < NS; k++, next_drain++) if (drain(e->slot[next_drain % NS]) != RC_OK) return RC_FAIL;
            if (failed_unit < 0 && rc2 == RC_OK && seq_unit >= 0) failed_unit = seq_unit;
            if (failed_unit >= 0 && units[(size_t)failed_unit].speculative) {
                /* the magic-scan boundary was wrong (or the member is corrupt): decode it alone,
                 * let the device find its end, check the footer found there, then resume parsing */
                ParsedMember u = units[(size_t)failed_unit];
                const uint64_t rest = c->src_len - u.m.src_off;
                const uint64_t capl = std::min<uint64_t>(std::min<uint64_t>(c->dst_cap - cur_out, rest * 1032 + 64), 0xfffffff0ull);
                QzbMemberResult r; const uint8_t *staged = nullptr;
                if (rest > 0xfffffff0ull) return RC_FAIL;
                if (run_single(u, (uint32_t)rest, (uint32_t)capl, &r, &staged) != RC_OK) return RC_FAIL;
                int st = status_rc(r, false);
                if (st == RC_OK && !r.saw_final) st = RC_DATA_ERROR;
                if (st == RC_OK && (uint64_t)r.consumed + 8 > rest) st = RC_DATA_ERROR;
                if (st == RC_OK) {
                    const uint8_t *ftr = hsrc + u.m.src_off + r.consumed;
                    if (rd32(ftr) != r.cksum || rd32(ftr + 4) != r.produced) st = RC_DATA_ERROR;
                }
                if (st != RC_OK) { final_rc = st; break; }
                if (deliver(staged, r.produced) != RC_OK) return RC_FAIL;
                cur_in = u.m.src_off + r.consumed + 8; cur_out += r.produced; o->nmembers++;
                if (c->stop_at_first) break;
                continue;                                    /* re-parse from the corrected position */
            }
            if (rc2 != RC_OK) { final_rc = rc2; break; }
        } else {
            /* 4B / RAW / unsized frames: output sizes come back from the device.  Decode into
             * private staging regions, then place the results one after another. */
            Slot &s = e->slot[0];
            size_t i = 0;
            const uint64_t bin = e->tune.batch_bytes;
            while (i < units.size() && rc2 == RC_OK) {
                size_t j = i; uint64_t sin = 0, sout = 0;
                while (j < units.size()) {
                    const uint64_t ulen = units[j].hdr_len + (uint64_t)units[j].m.src_len + units[j].ftr_len;
                    if (j > i && (sin + ulen > bin || sout > (e->tune.batch_bytes << 2))) break;
                    sin += ulen; sout += align_up(units[j].m.dst_cap, 16); j++;
                }
                if (run(s, i, j - i, units[i].unit_start, sin, 0, sout, true) != RC_OK) return RC_FAIL;
                s.busy = false;
                CK(cudaEventSynchronize(s.ev_meta));
                float ms = 0; cudaEventElapsedTime(&ms, s.ev_k0, s.ev_k1); o->kernel_ms += ms;
                std::vector<QzbMemberResult> res((const QzbMemberResult *)s.h_results.p, (const QzbMemberResult *)s.h_results.p + (j - i));
                std::vector<QzbMember> hm((const QzbMember *)s.h_members.p, (const QzbMember *)s.h_members.p + (j - i));
                const std::vector<ParsedMember> saved(units.begin() + (long)i, units.begin() + (long)j);
                for (size_t k = 0; k < saved.size(); k++) {
                    const ParsedMember &u = saved[k];
                    const uint8_t *from = (const uint8_t *)s.d_out.p + hm[k].dst_off;
                    QzbMemberResult r = res[k];
                    int st = status_rc(r, false);
                    if (st == RC_BUF_ERROR && (uint64_t)u.m.dst_cap < c->dst_cap - cur_out && c->fmt == QZB_FMT_4B) {
                        /* a block larger than hw_buff_sz (the reference hands these to zlib): retry it alone with room */
                        const uint64_t capl = std::min<uint64_t>(std::min<uint64_t>(c->dst_cap - cur_out, (uint64_t)u.m.src_len * 1032 + 64), 0xfffffff0ull);
                        const uint8_t *staged = nullptr;
                        if (run_single(u, u.m.src_len, (uint32_t)capl, &r, &staged) != RC_OK) return RC_FAIL;
                        units = saved;                      /* run_single reused the vector */
                        from = staged; st = status_rc(r, false);
                        if (st == RC_OK) {
                            if (deliver(from, r.produced) != RC_OK) return RC_FAIL;
                            cur_out += r.produced; cur_in = u.unit_start + u.hdr_len + u.m.src_len + u.ftr_len; o->nmembers++;
                            /* staging of the batch is gone: re-parse what follows */
                            rc2 = RC_OK; i = units.size(); j = i; goto reparse;
                        }
                    }
                    if (st == RC_OK && r.produced > c->dst_cap - cur_out) st = RC_BUF_ERROR;
                    if (st != RC_OK && c->fmt == QZB_FMT_RAW && r.safe_produced && r.safe_consumed && r.safe_produced <= c->dst_cap - cur_out) {
                        /* a raw stream that ran out of input (or room) behind a flush marker: everything up to that marker is
                         * delivered and the call reports how far it got, so a stream caller can go on from there with more
                         * input (reference: the piecemeal path of src/qatzip_stream.c:599-749) */
                        if (deliver(from, r.safe_produced) != RC_OK) return RC_FAIL;
                        cur_out += r.safe_produced; cur_in = u.unit_start + r.safe_consumed; o->nmembers++;
                    }
                    if (st != RC_OK) { rc2 = st; break; }
                    if (deliver(from, r.produced) != RC_OK) return RC_FAIL;
                    cur_out += r.produced;
                    cur_in = u.unit_start + u.hdr_len + ((u.m.exact_len & 1) ? u.m.src_len : r.consumed) + u.ftr_len;
                    o->nmembers++;
                }
                i = j;
            }
            if (rc2 != RC_OK) { final_rc = rc2; break; }
        }
        if (c->stop_at_first) break;
        if (parse_rc != RC_OK) { final_rc = parse_rc; break; }
reparse:;
    }
    o->consumed = cur_in; o->produced = cur_out;
    o->end_of_stream = (o->nmembers > 0) ? 1 : 0;
    return final_rc;
}
