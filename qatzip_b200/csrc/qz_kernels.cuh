/* qz_kernels.cuh -- launch-side descriptors shared by the kernels (qz_*.cu) and the host engine.
 * Everything here is plain data; no CUDA types leak above qz_device.h. */
#ifndef QZ_KERNELS_CUH
#define QZ_KERNELS_CUH
#include <stdint.h>
#include "qz_hd.h"

/* Compress side.  A *chunk* is the reference's unit (hw_buff_sz bytes of input -> one gzip
 * member / 4B block / LZ4 frame, reference src/qatzip.c:1513-1594).  A *piece* is the device's
 * unit: PIECE bytes of a chunk compressed by one warp into a byte-aligned run of deflate blocks
 * (or one LZ4 block).  Pieces of a chunk are laid end to end by the framing kernel. */
/* 32-bit words of token scratch per resident warp.  Deflate: 16-bit slots, at most one per input byte, the end-of-block slot,
 * and room for the 16-byte loads of the emit pass to run past the last one.  LZ4: two words per match. */
#define QZB_TOK_STRIDE(piece) ((piece) / 2 + 32)

struct QzbCompressJob {
    const uint8_t *src;          /* device, batch input */
    uint64_t src_len;
    uint32_t chunk_sz;           /* hw_buff_sz */
    uint32_t piece_log2;         /* 13 or 14 */
    uint32_t pieces_per_chunk;   /* ceil(chunk_sz / PIECE) */
    uint32_t nchunks, npieces;
    int32_t fmt;                 /* QzbFormat */
    int32_t last;                /* RAW: set BFINAL on the last chunk of the batch */
    int32_t static_huffman;      /* QZ_STATIC_HDR sessions: fixed codes only */
    uint8_t *slots;              /* npieces * slot_stride bytes, 16-byte aligned stride */
    uint32_t slot_stride;
    uint32_t *piece_len;         /* [npieces] bytes produced per piece */
    uint32_t *piece_crc;         /* [npieces] CRC-32 of the piece's input (zlib format: packed Adler sums) */
    uint32_t *tok_scratch;       /* [resident warps * QZB_TOK_STRIDE(PIECE)] token scratch */
    uint32_t *ticket;            /* dynamic piece counter (zeroed before launch) */
    /* framing */
    uint8_t *dst;                /* device, batch output */
    uint64_t dst_cap;
    uint32_t *chunk_total;       /* [nchunks] header + payload + footer bytes */
    uint64_t *chunk_off;         /* [nchunks + 1] exclusive prefix of chunk_total */
    uint32_t *chunk_cksum;       /* [nchunks] CRC-32, Adler-32 (zlib) or XXH32 of the chunk's input */
    uint32_t ngroups;            /* window kernels only: windows of 8 pieces (64 KiB) over all chunks (0 = per-piece kernel) */
    uint32_t tent;               /* window kernels: entries of a warp's hash table */
};

/* Decompress side: one unit = one gzip member / 4B block / raw stream / LZ4 frame. */
struct QzbMember {
    uint64_t src_off;            /* payload start (after the header) within the batch input */
    uint32_t src_len;            /* payload bytes available (exact when known, else upper bound) */
    uint32_t exact_len;          /* 1: payload must end exactly at src_len */
    uint64_t dst_off;            /* where the output goes */
    uint32_t dst_cap;            /* expected (exact_out=1) or maximum output bytes */
    uint32_t exact_out;          /* 1: produced must equal dst_cap (ISIZE / content size) */
    uint32_t expect_cksum;       /* footer CRC-32 / XXH32 */
    uint32_t check_cksum;        /* 1: verify it */
};
struct QzbMemberResult {
    uint32_t status;             /* QzbStatus */
    uint32_t consumed;           /* payload bytes consumed */
    uint32_t produced;
    uint32_t cksum;              /* CRC-32 / XXH32 of the output */
    uint32_t saw_final;          /* deflate: BFINAL block seen */
    uint32_t safe_consumed;      /* deflate: input and output position behind the last stored block (byte-aligned: every flush marker */
    uint32_t safe_produced;      /*   is one), where a decode that ran out of input or room can be picked up again; 0 if none */
    uint32_t pad;
};
struct QzbDecompressJob {
    const uint8_t *src;
    uint8_t *dst;
    const QzbMember *members;
    QzbMemberResult *results;
    uint32_t nmembers;
    int32_t fmt;
    uint32_t *ticket;
    int32_t size_only;           /* deflate: decode lengths only -- no output written, no checksum (member discovery) */
    const uint32_t *order;       /* deflate: ticket -> member index, largest payload first (NULL: in index order) */
};
#endif
