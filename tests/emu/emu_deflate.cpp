/* emu_deflate.cpp -- TEST INFRASTRUCTURE: the deflate piece kernel and the framing kernels of
 * qatzip_b200/csrc/qz_deflate.cu, compiled by g++ against the SIMT emulator (warp_emu.h) and driven the
 * way qz_engine.cu's enqueue_compress() drives them on the device.  Lets the CPU suite check the kernels'
 * warp logic (any geometry) without a GPU.  Nothing here is linked into libqatzip.so. */
#include "warp_emu.h"
#include "../../qatzip_b200/csrc/qz_deflate.cu"
#include <vector>

struct EmuCompressBuffers {
    std::vector<uint8_t> slots, meta;
    std::vector<uint32_t> tok;
    size_t slots_used = 0, meta_used = 0;      /* bytes the kernels may touch; the rest is canary */
};
static const size_t EMU_GUARD = 4096;
/* the canaries behind the slot and meta buffers are intact (a kernel that writes past its scratch shows up here) */
static bool emu_canaries_ok(const EmuCompressBuffers &b)
{
    for (size_t i = b.slots_used; i < b.slots.size(); i++) if (b.slots[i] != 0xEE) return false;
    for (size_t i = b.meta_used; i < b.meta.size(); i++) if (b.meta[i] != 0xEE) return false;
    return true;
}

extern "C" EmuCompressBuffers *emu_buffers_new(void) { return new EmuCompressBuffers(); }
extern "C" void emu_buffers_free(EmuCompressBuffers *b) { delete b; }

static size_t up16(size_t v) { return (v + 15) & ~(size_t)15; }

/* fills job's geometry + scratch pointers for `len` bytes of input */
extern "C" void emu_job_setup(QzbCompressJob *job, EmuCompressBuffers *b, int fmt, const uint8_t *src, uint64_t len, uint32_t chunk_sz, int last,
                              int static_huffman, int piece_log2, int resident_warps, uint8_t *dst, uint64_t cap)
{
    const uint32_t PIECE = 1u << piece_log2;
    memset(job, 0, sizeof *job);
    job->src = src; job->src_len = len; job->chunk_sz = chunk_sz; job->piece_log2 = (uint32_t)piece_log2;
    job->pieces_per_chunk = (chunk_sz + PIECE - 1) / PIECE;
    job->nchunks = len ? (uint32_t)((len + chunk_sz - 1) / chunk_sz) : 1u;
    const uint64_t last_len = len - (uint64_t)(job->nchunks - 1) * chunk_sz;
    const uint32_t last_pieces = last_len ? (uint32_t)((last_len + PIECE - 1) / PIECE) : 1u;
    job->npieces = (job->nchunks - 1) * job->pieces_per_chunk + last_pieces;
    job->fmt = fmt; job->last = last; job->static_huffman = static_huffman;
    job->slot_stride = PIECE + 64;
    b->slots_used = (size_t)job->npieces * job->slot_stride + 64;
    b->slots.assign(b->slots_used + EMU_GUARD, 0xEE);
    size_t o = 0;
    const size_t o_off = o; o += up16((size_t)(job->nchunks + 1) * 8);
    const size_t o_ck = o; o += up16((size_t)job->nchunks * 4);
    const size_t o_tot = o; o += up16((size_t)job->nchunks * 4);
    const size_t o_plen = o; o += up16((size_t)job->npieces * 4);
    const size_t o_pcrc = o; o += up16((size_t)job->npieces * 4);
    const size_t o_ticket = o; o += 16;
    b->meta_used = o;
    b->meta.assign(o + EMU_GUARD, 0xEE);
    uint8_t *m = b->meta.data();
    memset(m + o_ticket, 0, 16);
    job->slots = b->slots.data();
    job->piece_len = (uint32_t *)(m + o_plen); job->piece_crc = (uint32_t *)(m + o_pcrc);
    job->chunk_total = (uint32_t *)(m + o_tot); job->chunk_cksum = (uint32_t *)(m + o_ck);
    job->chunk_off = (uint64_t *)(m + o_off); job->ticket = (uint32_t *)(m + o_ticket);
    b->tok.assign((size_t)resident_warps * QZB_TOK_STRIDE(PIECE), 0xEEEEEEEEu);
    job->tok_scratch = b->tok.data();
    job->dst = dst; job->dst_cap = cap;
}

/* sizes -> scan -> frame, as qzb_launch_frame does; returns the bytes laid out in dst (chunks that do not fit are left out) */
extern "C" long emu_frame(const QzbCompressJob *jobp, uint32_t *chunk_cksum_out)
{
    const QzbCompressJob job = *jobp;
    emu::launch((job.nchunks + 255) / 256, 256, 0, [&] { qzb_chunk_sizes_kernel(job); });
    emu::launch(1, 1024, 0, [&] { qzb_scan_kernel(job.chunk_total, job.chunk_off, job.nchunks); });
    emu::launch(job.nchunks, QZ_FRAME_WARPS * 32, 0, [&] { qzb_frame_kernel(job); });
    uint32_t fit = 0;
    while (fit < job.nchunks && job.chunk_off[fit + 1] <= job.dst_cap) fit++;
    if (chunk_cksum_out) memcpy(chunk_cksum_out, job.chunk_cksum, (size_t)fit * 4);
    return (long)job.chunk_off[fit];
}

/* One batch through the deflate kernels.  Geometry is the caller's.  Returns bytes produced, -1 for an unsupported geometry,
 * -2 when a kernel wrote past its scratch.
 * window == 0: per-piece kernel (piece size, hash bits, warps and piece buffers per CTA, CTAs);
 * window != 0: window kernel (one deflate block per 64 KiB window; window == 2: two-entry hash buckets): CTAs of 32 warps, tables of `hb` entries each when
 *              hb >= 256, else 2^hb (warps and nbuf are ignored). */
extern "C" long emu_deflate_compress(int fmt, const uint8_t *src, uint64_t len, uint32_t chunk_sz, int last, int static_huffman,
                                     int piece_log2, int hb, int warps, int nbuf, int grid, uint8_t *dst, uint64_t cap, uint32_t *chunk_cksum_out, int window)
{
    if (grid < 1 || (!window && (warps < 1 || warps > 32 || nbuf < 1 || nbuf > warps))) return -1;
    QzbCompressJob job; EmuCompressBuffers b;
    emu_job_setup(&job, &b, fmt, src, len, chunk_sz, last, static_huffman, piece_log2, grid * warps, dst, cap);
    size_t smem;
    std::function<void()> body;
    if (window) {
        if (!len || job.pieces_per_chunk % QZ_WINDOW_PIECES || piece_log2 != 13) return -1;
        const uint64_t last_len = len - (uint64_t)(job.nchunks - 1) * chunk_sz;
        job.ngroups = (job.nchunks - 1) * (job.pieces_per_chunk / QZ_WINDOW_PIECES) + (uint32_t)((last_len + QZ_WINDOW - 1) / QZ_WINDOW);
        job.tent = hb >= 256 ? (uint32_t)hb : 1u << hb;
        smem = (size_t)window_unit_bytes(job.tent) + 2 * sizeof(BlockCoder);
        if (smem > 227 * 1024) return -1;
        b.tok.assign((size_t)grid * QZW_MATCHERS * QZW_TOK_WORDS, 0xEEEEEEEEu);
        job.tok_scratch = b.tok.data();
        if (window == 2) emu::launch((unsigned)grid, QZW_WARPS * 32, smem, [&] { qzb_deflate_window_kernel<2>(job); });
        else emu::launch((unsigned)grid, QZW_WARPS * 32, smem, [&] { qzb_deflate_window_kernel<1>(job); });
        const long n = emu_frame(&job, chunk_cksum_out);
        return emu_canaries_ok(b) ? n : -2;
    }
    if (piece_log2 == 13 && hb == 11) { smem = sizeof(WarpPriv<11>) * warps + sizeof(PieceBuf<13>) * nbuf; body = [&] { qzb_deflate_pieces_kernel<13, 11>(job, nbuf); }; }
    else if (piece_log2 == 13 && hb == 12) { smem = sizeof(WarpPriv<12>) * warps + sizeof(PieceBuf<13>) * nbuf; body = [&] { qzb_deflate_pieces_kernel<13, 12>(job, nbuf); }; }
    else if (piece_log2 == 14 && hb == 12) { smem = sizeof(WarpPriv<12>) * warps + sizeof(PieceBuf<14>) * nbuf; body = [&] { qzb_deflate_pieces_kernel<14, 12>(job, nbuf); }; }
    else if (piece_log2 == 14 && hb == 13) { smem = sizeof(WarpPriv<13>) * warps + sizeof(PieceBuf<14>) * nbuf; body = [&] { qzb_deflate_pieces_kernel<14, 13>(job, nbuf); }; }
    else return -1;
    emu::launch((unsigned)grid, (unsigned)warps * 32, smem, body);
    const long n = emu_frame(&job, chunk_cksum_out);
    return emu_canaries_ok(b) ? n : -2;
}

extern "C" unsigned long long emu_collectives(void) { return emu::collectives(); }
