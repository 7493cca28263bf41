/* placeholder until the LZ4 kernels land (next commit) */
#include <cuda_runtime.h>
#include "qz_kernels.cuh"
extern "C" size_t qzb_lz4_smem_bytes(int piece_log2, int warps) { return ((size_t)(1 << piece_log2) + 32 + 4096 + 64) * (size_t)warps; }
extern "C" cudaError_t qzb_launch_lz4_compress(const QzbCompressJob *, int, int, cudaStream_t) { return cudaErrorNotSupported; }
extern "C" cudaError_t qzb_launch_lz4_decompress(const QzbDecompressJob *, int, cudaStream_t) { return cudaErrorNotSupported; }
