/* emu_inflate.cpp -- TEST INFRASTRUCTURE: qzb_inflate_kernel (qatzip_b200/csrc/qz_inflate.cu) compiled by g++ against
 * the SIMT emulator and launched over a caller-made member table, the way qz_engine.cu launches it.  Not part of libqatzip.so. */
#include "warp_emu.h"
#include "../../qatzip_b200/csrc/qz_inflate.cu"

extern "C" int emu_inflate(int fmt, const uint8_t *src, uint8_t *dst, const QzbMember *members, QzbMemberResult *results, uint32_t nmembers,
                           int size_only, int grid)
{
    uint32_t ticket[4] = { 0, 0, 0, 0 };
    QzbDecompressJob job; memset(&job, 0, sizeof job);
    job.src = src; job.dst = dst; job.members = members; job.results = results; job.nmembers = nmembers; job.fmt = fmt;
    job.ticket = ticket; job.size_only = size_only;
    /* decoders per warp: QZ_EMU_INFLATE_DPW (1, 2, 4, 8), default 4 as in the product */
    const char *e = getenv("QZ_EMU_INFLATE_DPW");
    const int dpw = e ? atoi(e) : 4;
    const unsigned g = (unsigned)(grid < 1 ? 1 : grid);
    if (dpw == 1) emu::launch(g, 256, sizeof(InflWarpSmem) * 8, [&] { qzb_inflate_kernel<1>(job); });
    else if (dpw == 2) emu::launch(g, 256, sizeof(InflWarpSmem) * 16, [&] { qzb_inflate_kernel<2>(job); });
    else if (dpw == 8) emu::launch(g, 128, sizeof(InflWarpSmem) * 32, [&] { qzb_inflate_kernel<8>(job); });
    else emu::launch(g, 256, sizeof(InflWarpSmem) * 32, [&] { qzb_inflate_kernel<4>(job); });
    return 0;
}

/* qzb_gzip_scan_kernel over src[lo, n): offsets (relative to lo) of plausible gzip member starts, unordered; returns how many
 * were found (more than cap: the list is cut, the count is not) */
extern "C" int emu_gzip_scan(const uint8_t *src, uint64_t lo, uint64_t n, uint32_t *list, uint32_t cap, int grid)
{
    uint32_t count = 0;
    emu::launch((unsigned)(grid < 1 ? 1 : grid), 256, 0, [&] { qzb_gzip_scan_kernel(src, lo, n, list, cap, &count); });
    return (int)count;
}
