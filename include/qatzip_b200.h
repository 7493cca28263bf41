/* qatzip_b200.h -- extensions of the qatzip.h ABI that only make sense on a GPU back end.
 * Nothing in the reference corresponds to these (its device is reached through host buffers
 * only); they exist so that callers whose data already lives in HBM, and bench.py's
 * roofline leg, can run the same kernels without the PCIe copies.  Plain C, no CUDA types. */
#ifndef QATZIP_B200_H
#define QATZIP_B200_H
#include <stddef.h>
#include <stdint.h>
#include "qatzip.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct QzB200Stats_S {
    double kernel_ms;            /* device time of all kernels of the last call (CUDA events on the launch stream) */
    double codec_ms;             /* of which: the piece kernel (deflate / LZ4 compress), summed over its launches */
    uint64_t codec_launches;     /* launches of the piece kernel in the last call */
    uint64_t kernel_launches;    /* kernels launched by the last call */
    uint64_t units;              /* chunks compressed / members decoded by the last call */
    int device;                  /* CUDA device ordinal the session runs on */
    int piece_log2, hash_bits;   /* compressor geometry in use */
    double h2d_ms, d2h_ms;       /* host-buffer compress calls: summed copy times of the last call (overlapping the kernels) */
    int group_blocks;            /* 1: deflate sessions of this hw_buff_sz get one block per 64 KiB window (window kernel), 0: one per 8 KiB piece */
    int devices;                 /* GPUs a host-buffer compress call of this session is spread over (QZB200_DEVICES; 1 by default) */
} QzB200Stats_T;

/* qzCompress / qzDecompress with src and dest in device memory of the session's GPU.
 * Same return codes and partial-progress rules as the host entry points; lengths are 64-bit.
 * For decompress `h_src_view` is a host copy of the compressed bytes (member headers are
 * walked on the host, as checkHeader does in the reference: src/qatzip_utils.c:1232). */
int qzb200CompressDevice(QzSession_T *sess, const void *d_src, uint64_t src_len, void *d_dest, uint64_t dest_cap,
                         unsigned int last, uint64_t *consumed, uint64_t *produced, unsigned long *crc);
int qzb200DecompressDevice(QzSession_T *sess, const void *d_src, const void *h_src_view, uint64_t src_len,
                           void *d_dest, uint64_t dest_cap, uint64_t *consumed, uint64_t *produced);
/* statistics of the most recent data call on this session (host or device entry point) */
int qzb200GetStats(QzSession_T *sess, QzB200Stats_T *stats);
/* plain device-memory helpers (cudaMalloc / cudaFree / cudaMemcpy on the session's GPU) so that a
 * C caller or a ctypes harness can stage data in HBM without linking the CUDA runtime itself */
void *qzb200DeviceAlloc(uint64_t bytes);
void qzb200DeviceFree(void *d_ptr);
int qzb200CopyToDevice(void *d_dst, const void *h_src, uint64_t bytes);
int qzb200CopyToHost(void *h_dst, const void *d_src, uint64_t bytes);
/* number of CUDA devices the library can use, and the one this process would pick (QZB200_DEVICE, else LOCAL_RANK, else 0).
 * QZB200_DEVICES = "all" | count | list makes one process use several: sessions take them in turn as their primary device,
 * and every host-buffer qzCompress call deals its batches over all of them (output and checksums stitched in order). */
int qzb200DeviceCount(void);
int qzb200DefaultDevice(void);

#ifdef __cplusplus
}
#endif
#endif
