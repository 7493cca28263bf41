/* qz_engine.h -- host-side chunk engine: the B200 replacement for the reference's
 * submit / poll / stitch loops (reference src/qatzip.c:1483-1764 compress, :2103-2404
 * decompress).  A QzbEngine belongs to one session; it owns CUDA streams, device buffers and
 * pinned bounce buffers on one GPU and pipelines batches of chunks H2D -> kernels -> D2H. */
#ifndef QZ_ENGINE_H
#define QZ_ENGINE_H
#include <stdint.h>
#include <stddef.h>
#include "qz_hd.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct QzbEngine QzbEngine;

/* process-wide runtime: returns the number of usable CUDA devices (0 if none / no driver) */
int qzb_runtime_devices(void);
/* which device this process uses by default: QZB200_DEVICE, else LOCAL_RANK, else 0 */
int qzb_runtime_default_device(void);

/* devices host-buffer compress calls may be spread over (QZB200_DEVICES; default: the one default device) */
int qzb_runtime_device_list(int *out, int cap);
QzbEngine *qzb_engine_create(int device);
int qzb_engine_device_count(const QzbEngine *e);
int qzb_engine_primary_device(const QzbEngine *e);
void qzb_engine_destroy(QzbEngine *e);

typedef struct QzbCompressCall {
    int fmt;                 /* QzbFormat */
    int level;
    int static_huffman;
    int last;
    uint32_t chunk_sz;       /* hw_buff_sz */
    const uint8_t *src; uint64_t src_len;
    uint8_t *dst; uint64_t dst_cap;
    int src_device, dst_device;     /* 1: pointer is device memory on the engine's GPU */
    int src_pinned, dst_pinned;     /* 1: host pointer is page-locked (DMA without staging) */
    int want_crc;
    uint32_t crc_in;                /* running CRC (0 restarts), deflate formats */
} QzbCompressCall;
typedef struct QzbCompressOut {
    uint64_t consumed, produced;
    uint32_t crc;
    uint32_t nchunks;
    double kernel_ms;               /* device time of codec + framing kernels (CUDA events), all batches */
    double codec_ms;                /* device time of the piece kernel alone (deflate / lz4) */
    uint64_t codec_launches;        /* how many times the piece kernel was launched */
    uint64_t kernel_launches;
    double h2d_ms, d2h_ms;          /* host-buffer calls: summed copy times (CUDA events), they overlap with kernels */
} QzbCompressOut;
/* returns a qatzip.h return code (QZ_OK, QZ_BUF_ERROR with partial progress, QZ_FAIL) */
int qzb_engine_compress(QzbEngine *e, const QzbCompressCall *c, QzbCompressOut *o);

typedef struct QzbDecompressCall {
    int fmt;
    uint32_t chunk_sz;              /* hw_buff_sz: cap for formats that do not carry sizes (4B) */
    const uint8_t *src; uint64_t src_len;
    uint8_t *dst; uint64_t dst_cap;
    int src_device, dst_device, src_pinned, dst_pinned;
    const uint8_t *src_host_view;   /* when src_device: a host copy of the same bytes for header parsing */
    int stop_at_first;              /* decode exactly one member (stop_decompression_stream_end) */
} QzbDecompressCall;
typedef struct QzbDecompressOut {
    uint64_t consumed, produced;
    uint32_t nmembers;
    int end_of_stream;
    double kernel_ms;
    uint64_t kernel_launches;
} QzbDecompressOut;
int qzb_engine_decompress(QzbEngine *e, const QzbDecompressCall *c, QzbDecompressOut *o);

/* pinned-page registry shared with qzMalloc (reference src/qatzip_mem.c:102 qzMemFindAddr) */
void *qzb_pinned_alloc(size_t sz);
int qzb_pinned_free(void *p);          /* 1 if p was one of ours */
int qzb_pinned_contains(const void *p, size_t len);

/* tuning knobs (environment: QZB200_PIECE_LOG2, QZB200_HASH_BITS, QZB200_BATCH_MB, QZB200_WARPS) */
typedef struct QzbTuning { int piece_log2, hash_bits, warps_per_cta, buffers_per_cta, inflate_dpw; size_t batch_bytes, first_batch_bytes, zlib_window_bytes, inflate_batch_bytes; int taper; int window, window_tent, lz4_warps; } QzbTuning;
void qzb_get_tuning(QzbTuning *t);

/* raw device memory helpers for callers that keep data in HBM (bench, tests) */
void *qzb_device_alloc(int device, size_t n);
void qzb_device_free(int device, void *p);
int qzb_device_copy(int device, void *dst, const void *src, size_t n, int to_device);

#ifdef __cplusplus
}
#endif
#endif
