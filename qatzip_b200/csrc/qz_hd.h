/* qz_hd.h -- macros that let the scalar pieces of the codec (tables, Huffman construction,
 * header coding, inflate decode loop, CRC-32, xxHash32) compile both as CUDA device code and
 * as plain host C++ (for CPU unit tests of the exact same source). */
#ifndef QZ_HD_H
#define QZ_HD_H
#include <stdint.h>
#include <stddef.h>
#if defined(__CUDACC__)
#define QZ_HD __host__ __device__ __forceinline__
#define QZ_HD_SERIAL static __host__ __device__ __noinline__   /* long single-lane routines: one copy, out of the hot I-cache path */
#else
#define QZ_HD static inline
#define QZ_HD_SERIAL static
#endif

/* The scalar routines take plain pointers; in the kernels those always point into shared memory.  Saying so lets the
 * compiler use shared-memory loads and stores (LDS/STS) instead of generic ones inside out-of-line routines. */
#if defined(__CUDA_ARCH__)
#define QZ_ASSUME_SHARED(p) __builtin_assume(__isShared(p))
#else
#define QZ_ASSUME_SHARED(p) do { } while (0)
#endif

/* wire/data formats handled by the kernels (superset of QzDataFormat_T: LZ4 is a session type) */
enum QzbFormat { QZB_FMT_4B = 0, QZB_FMT_GZIP = 1, QZB_FMT_GZIP_EXT = 2, QZB_FMT_RAW = 3, QZB_FMT_LZ4 = 4, QZB_FMT_ZLIB = 5 };

/* per-unit status words written by kernels */
enum QzbStatus { QZB_ST_OK = 0, QZB_ST_DATA_ERROR = 1, QZB_ST_OUT_FULL = 2, QZB_ST_IN_TRUNC = 3, QZB_ST_CKSUM = 4, QZB_ST_SIZE = 5 };

#endif
