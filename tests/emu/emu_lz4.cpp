/* emu_lz4.cpp -- TEST INFRASTRUCTURE: the LZ4 kernels (qatzip_b200/csrc/qz_lz4.cu) compiled by g++ against the SIMT
 * emulator; framing comes from emu_deflate.cpp (qzb_frame_kernel).  Not part of libqatzip.so. */
#include "warp_emu.h"
#include "../../qatzip_b200/csrc/qz_lz4.cu"
#include <vector>

struct EmuCompressBuffers;
extern "C" void emu_job_setup(QzbCompressJob *job, EmuCompressBuffers *b, int fmt, const uint8_t *src, uint64_t len, uint32_t chunk_sz, int last,
                              int static_huffman, int piece_log2, int resident_warps, uint8_t *dst, uint64_t cap);
extern "C" long emu_frame(const QzbCompressJob *job, uint32_t *chunk_cksum_out);
extern "C" EmuCompressBuffers *emu_buffers_new(void);
extern "C" void emu_buffers_free(EmuCompressBuffers *b);

extern "C" long emu_lz4_compress(const uint8_t *src, uint64_t len, uint32_t chunk_sz, int piece_log2, int warps, int grid, uint8_t *dst, uint64_t cap,
                                 uint32_t *chunk_cksum_out)
{
    if (warps < 1 || warps > 16 || grid < 1 || (piece_log2 != 13 && piece_log2 != 14)) return -1;
    QzbCompressJob job; EmuCompressBuffers *b = emu_buffers_new();
    emu_job_setup(&job, b, QZB_FMT_LZ4, src, len, chunk_sz, 1, 0, piece_log2, grid * warps, dst, cap);
    if (piece_log2 == 13) emu::launch((unsigned)grid, (unsigned)warps * 32, sizeof(Lz4WarpSmem<13>) * warps, [&] { qzb_lz4_pieces_kernel<13>(job); });
    else emu::launch((unsigned)grid, (unsigned)warps * 32, sizeof(Lz4WarpSmem<14>) * warps, [&] { qzb_lz4_pieces_kernel<14>(job); });
    emu::launch((job.nchunks * 4 + 255) / 256, 256, 0, [&] { qzb_xxh32_chunks_kernel(job); });
    const long n = emu_frame(&job, chunk_cksum_out);
    emu_buffers_free(b);
    return n;
}

/* window kernel: one LZ4 block per 64 KiB window, `nw` (12 or 16) warps, tables of `tent` entries */
extern "C" long emu_lz4_window(const uint8_t *src, uint64_t len, uint32_t chunk_sz, int tent, int nw, int grid, uint8_t *dst, uint64_t cap, uint32_t *chunk_cksum_out)
{
    if (grid < 1 || !len || chunk_sz % 65536 || (nw != 12 && nw != 16)) return -1;
    QzbCompressJob job; EmuCompressBuffers *b = emu_buffers_new();
    emu_job_setup(&job, b, QZB_FMT_LZ4, src, len, chunk_sz, 1, 0, 13, 1, dst, cap);
    const uint64_t last_len = len - (uint64_t)(job.nchunks - 1) * chunk_sz;
    job.ngroups = (job.nchunks - 1) * (job.pieces_per_chunk / 8) + (uint32_t)((last_len + 65535) / 65536);
    job.tent = (uint32_t)tent;
    std::vector<uint32_t> tok((size_t)grid * nw * QZB_TOK_STRIDE(lz4_sub_bytes(nw)), 0xEEEEEEEEu);
    job.tok_scratch = tok.data();
    const size_t smem = lz4_window_smem(job.tent, nw);
    if (smem > 227 * 1024) { emu_buffers_free(b); return -1; }
    if (nw == 12) emu::launch((unsigned)grid, 12 * 32, smem, [&] { qzb_lz4_window_kernel<12>(job); });
    else emu::launch((unsigned)grid, 16 * 32, smem, [&] { qzb_lz4_window_kernel<16>(job); });
    emu::launch((job.nchunks * 4 + 255) / 256, 256, 0, [&] { qzb_xxh32_chunks_kernel(job); });
    const long n = emu_frame(&job, chunk_cksum_out);
    emu_buffers_free(b);
    return n;
}

extern "C" int emu_lz4_decompress(const uint8_t *src, uint8_t *dst, const QzbMember *members, QzbMemberResult *results, uint32_t nmembers, int grid)
{
    uint32_t ticket[4] = { 0, 0, 0, 0 };
    QzbDecompressJob job; memset(&job, 0, sizeof job);
    job.src = src; job.dst = dst; job.members = members; job.results = results; job.nmembers = nmembers; job.fmt = QZB_FMT_LZ4; job.ticket = ticket;
    emu::launch((unsigned)(grid < 1 ? 1 : grid), 256, 0, [&] { qzb_lz4_decompress_kernel(job); });
    return 0;
}
