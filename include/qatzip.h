/* qatzip.h -- C ABI of the B200-native chunked codec ("qatzip_b200").
 *
 * This header is the drop-in boundary.  Every type, constant and prototype below is
 * ABI-identical to intel/QATzip's public header (reference include/qatzip.h, API version 2.5,
 * library 1.3.1) so that a program compiled against the reference header links and runs
 * against libqatzip.so from this repository unchanged.  The declarations were re-written
 * here (no reference text is reproduced); each block cites the reference lines it mirrors.
 * Struct layouts are checked in tests/test_abi.py against the sizes/offsets probed from the
 * reference header.
 *
 * What sits behind the ABI is different: there is no QAT device and no CPU fall-back.  The
 * "hardware" is one NVIDIA B200 per process (sm_100a CUDA kernels); `sw_backup` is accepted
 * and ignored; if no CUDA device is usable qzInit() returns QZ_NOSW_NO_HW.
 */
#ifndef _QATZIP_H
#define _QATZIP_H

#ifdef __cplusplus
extern "C" {
#endif

#include <string.h>
#include <stdint.h>

/* reference include/qatzip.h:71-89 */
#define QATZIP_API_VERSION_NUM_MAJOR (2)
#define QATZIP_API_VERSION_NUM_MINOR (5)
#define QATZIP_API_VERSION (QATZIP_API_VERSION_NUM_MAJOR * 10000 + QATZIP_API_VERSION_NUM_MINOR * 100)
#define QATZIP_API

/* ---- enumerations: reference include/qatzip.h:200-300 ---- */
typedef enum QzHuffmanHdr_E { QZ_DYNAMIC_HDR = 0, QZ_STATIC_HDR } QzHuffmanHdr_T;
typedef enum PinMem_E { COMMON_MEM = 0, PINNED_MEM } PinMem_T;
typedef enum QzDirection_E { QZ_DIR_COMPRESS = 0, QZ_DIR_DECOMPRESS, QZ_DIR_BOTH } QzDirection_T;
/* Wire format of the deflate family.  LZ4 frames are chosen by the session type
 * (qzSetupSessionLZ4), not by this enum. */
typedef enum QzDataFormat_E {
    QZ_DEFLATE_4B = 0,     /* LE32 payload length + raw deflate, per chunk */
    QZ_DEFLATE_GZIP,       /* RFC 1952 member per chunk, 10-byte header */
    QZ_DEFLATE_GZIP_EXT,   /* RFC 1952 member per chunk with the 'QZ' extra field (sizes) */
    QZ_DEFLATE_RAW,        /* bare RFC 1951 stream */
    QZ_FMT_NUM
} QzDataFormat_T;
typedef enum QzPollingMode_E { QZ_PERIODICAL_POLLING = 0, QZ_BUSY_POLLING } QzPollingMode_T;
typedef enum QzCrcType_E { QZ_CRC32 = 0, QZ_ADLER, NONE } QzCrcType_T;
typedef enum QzSoftwareComponentType_E {
    QZ_COMPONENT_FIRMWARE = 0, QZ_COMPONENT_KERNEL_DRIVER, QZ_COMPONENT_USER_DRIVER,
    QZ_COMPONENT_QATZIP_API, QZ_COMPONENT_SOFTWARE_PROVIDER
} QzSoftwareComponentType_T;

/* ---- return codes: reference include/qatzip.h:311-361 ---- */
#define QZ_OK (0)
#define QZ_DUPLICATE (1)
#define QZ_FORCE_SW (2)
#define QZ_PARAMS (-1)
#define QZ_FAIL (-2)
#define QZ_BUF_ERROR (-3)
#define QZ_DATA_ERROR (-4)
#define QZ_TIMEOUT (-5)
#define QZ_INTEG (-100)
#define QZ_NO_HW (11)
#define QZ_NO_MDRV (12)
#define QZ_NO_INST_ATTACH (13)
#define QZ_LOW_MEM (14)
#define QZ_LOW_DEST_MEM (15)
#define QZ_UNSUPPORTED_FMT (16)
#define QZ_NONE (100)
#define QZ_NOSW_NO_HW (-101)
#define QZ_NOSW_NO_MDRV (-102)
#define QZ_NOSW_NO_INST_ATTACH (-103)
#define QZ_NOSW_LOW_MEM (-104)
#define QZ_NO_SW_AVAIL (-105)
#define QZ_NOSW_UNSUPPORTED_FMT (-116)
#define QZ_POST_PROCESS_ERROR (-117)
#define QZ_METADATA_OVERFLOW (-118)
#define QZ_OUT_OF_RANGE (-119)
#define QZ_NOT_SUPPORTED (-200)

/* algorithm tags: reference include/qatzip.h:363-372 */
#define QZ_MAX_ALGORITHMS ((int)255)
#define QZ_DEFLATE ((unsigned char)8)
#define QZ_LZ4 ((unsigned char)'4')
#define QZ_LZ4_BLOCK ((unsigned char)'B')
#define QZ_LZ4s ((unsigned char)'s')
#define QZ_ZSTD ((unsigned char)'Z')

#ifndef MIN
#define MIN(a,b) (((a)<(b))?(a):(b))
#endif
#define QZ_MEMCPY(dest, src, dest_sz, src_sz) memcpy((void *)(dest), (void *)(src), (size_t)MIN(dest_sz, src_sz))

/* ---- session parameter blocks: reference include/qatzip.h:440-570 ---- */
typedef int (*qzLZ4SCallbackFn)(void *external, const unsigned char *src, unsigned int *src_len,
                                unsigned char *dest, unsigned int *dest_len, int *ExtStatus);

typedef struct QzSessionParams_S {
    QzHuffmanHdr_T huffman_hdr;      /* dynamic (default) or static Huffman blocks */
    QzDirection_T direction;
    QzDataFormat_T data_fmt;
    unsigned int comp_lvl;           /* 1..9 */
    unsigned char comp_algorithm;    /* QZ_DEFLATE or QZ_LZ4 */
    unsigned int max_forks;
    unsigned char sw_backup;         /* accepted, ignored: there is no CPU path */
    unsigned int hw_buff_sz;         /* chunk size: power of two in [1 KiB, 512 KiB] */
    unsigned int strm_buff_sz;       /* stream staging size */
    unsigned int input_sz_thrshold;  /* accepted; small inputs still run on the GPU */
    unsigned int req_cnt_thrshold;
    unsigned int wait_cnt_thrshold;
#ifdef ERR_INJECTION
    void *fbError;
    void *fbErrorCurr;
#endif
} QzSessionParams_T;

typedef struct QzSessionParamsCommon_S {
    QzDirection_T direction;
    unsigned int comp_lvl;
    unsigned char comp_algorithm;
    unsigned int max_forks;
    unsigned char sw_backup;
    unsigned int hw_buff_sz;
    unsigned int strm_buff_sz;
    unsigned int input_sz_thrshold;
    unsigned int req_cnt_thrshold;
    unsigned int wait_cnt_thrshold;
    QzPollingMode_T polling_mode;
    unsigned int is_sensitive_mode;
#ifdef ERR_INJECTION
    void *fbError;
    void *fbErrorCurr;
#endif
} QzSessionParamsCommon_T;

typedef struct QzSessionParamsDeflate_S {
    QzSessionParamsCommon_T common_params;
    QzHuffmanHdr_T huffman_hdr;
    QzDataFormat_T data_fmt;
} QzSessionParamsDeflate_T;

typedef struct QzSessionParamsLZ4_S {
    QzSessionParamsCommon_T common_params;
} QzSessionParamsLZ4_T;

typedef struct QzSessionParamsLZ4S_S {
    QzSessionParamsCommon_T common_params;
    qzLZ4SCallbackFn qzCallback;
    void *qzCallback_external;
    unsigned int lz4s_mini_match;
} QzSessionParamsLZ4S_T;

typedef struct QzSessionParamsDeflateExt_S {
    QzSessionParamsDeflate_T deflate_params;
    unsigned char stop_decompression_stream_end;
    unsigned char zlib_format;
} QzSessionParamsDeflateExt_T;

/* defaults and limits: reference include/qatzip.h:573-612 */
#define QZ_HUFF_HDR_DEFAULT QZ_DYNAMIC_HDR
#define QZ_DIRECTION_DEFAULT QZ_DIR_BOTH
#define QZ_DATA_FORMAT_DEFAULT QZ_DEFLATE_GZIP_EXT
#define QZ_COMP_LEVEL_DEFAULT 1
#define QZ_COMP_ALGOL_DEFAULT QZ_DEFLATE
#define QZ_POLL_SLEEP_DEFAULT 10
#define QZ_MAX_FORK_DEFAULT 3
#define QZ_SW_BACKUP_DEFAULT 1
#define QZ_HW_BUFF_SZ (64*1024)
#define QZ_HW_BUFF_SZ_Gen3 (1*1024*1024)
#define QZ_HW_BUFF_MIN_SZ (1*1024)
#define QZ_HW_BUFF_MAX_SZ (512*1024)
#define QZ_HW_BUFF_MAX_SZ_Gen3 (2*1024*1024*1024U)
#define QZ_STRM_BUFF_SZ_DEFAULT QZ_HW_BUFF_SZ
#define QZ_STRM_BUFF_MIN_SZ (1*1024)
#define QZ_STRM_BUFF_MAX_SZ (2*1024*1024 - 5*1024)
#define QZ_COMP_THRESHOLD_DEFAULT 1024
#define QZ_COMP_THRESHOLD_MINIMUM 128
#define QZ_REQ_THRESHOLD_MINIMUM 1
#define QZ_REQ_THRESHOLD_MAXIMUM NUM_BUFF
#define QZ_REQ_THRESHOLD_DEFAULT QZ_REQ_THRESHOLD_MAXIMUM
#define QZ_WAIT_CNT_THRESHOLD_DEFAULT 8
#define QZ_DEFLATE_COMP_LVL_MINIMUM (1)
#define QZ_DEFLATE_COMP_LVL_MAXIMUM (9)
#define QZ_DEFLATE_COMP_LVL_MAXIMUM_Gen3 (12)
#define QZ_LZS_COMP_LVL_MINIMUM (1)
#define QZ_LZS_COMP_LVL_MAXIMUM (12)
#define QZ_AUTO_SELECT_NUMA_NODE (-1)

/* sw_backup / ext_rc bit helpers: reference include/qatzip.h:614-664 */
#define QZ_SW_BACKUP_BIT_POSITION (0)
#define QZ_SW_FORCESW_BIT_POSITION (1)
#define QZ_ENABLE_SOFTWARE_BACKUP(v) ((v) |= (1 << QZ_SW_BACKUP_BIT_POSITION))
#define QZ_ENABLE_SOFTWARE_ONLY_EXECUTION(v) ((v) |= (1 << QZ_SW_FORCESW_BIT_POSITION))
#define QZ_DISABLE_SOFTWARE_BACKUP(v) ((v) &= ~(1 << QZ_SW_BACKUP_BIT_POSITION))
#define QZ_DISABLE_SOFTWARE_ONLY_EXECUTION(v) ((v) &= ~(1 << QZ_SW_FORCESW_BIT_POSITION))
#define QZ_SW_EXECUTION_BIT (4)
#define QZ_SW_EXECUTION_MASK (1 << QZ_SW_EXECUTION_BIT)
#define QZ_SW_EXECUTION(ret, ext_rc) (!ret && (ext_rc & QZ_SW_EXECUTION_MASK))
#define QZ_TIMEOUT_BIT (8)
#define QZ_TIMEOUT_MASK (1 << QZ_TIMEOUT_BIT)
#define QZ_HW_TIMEOUT(ret, ext_rc) (!ret && (ext_rc & QZ_TIMEOUT_MASK))
#define QZ_POST_PROCESS_FAIL_BIT (10)
#define QZ_POST_PROCESS_FAIL_MASK (1 << QZ_POST_PROCESS_FAIL_BIT)
#define QZ_POST_PROCESS_FAIL(ret, ext_rc) (ret && (ext_rc & QZ_POST_PROCESS_FAIL_MASK))

/* caller-owned session handle (zero-initialise before first use): reference include/qatzip.h:676-687 */
typedef struct QzSession_S {
    signed long int hw_session_stat;
    int thd_sess_stat;
    void *internal;
    unsigned long total_in;
    unsigned long total_out;
} QzSession_T;

/* reference include/qatzip.h:689-760 */
typedef struct QzStatus_S {
    unsigned short int qat_hw_count;
    unsigned char qat_service_init;
    unsigned char qat_mem_drvr;
    unsigned char qat_instance_attach;
    unsigned long int memory_alloced;
    unsigned char using_huge_pages;
    signed long int hw_session_status;
    unsigned char algo_sw[QZ_MAX_ALGORITHMS];
    unsigned char algo_hw[QZ_MAX_ALGORITHMS];
} QzStatus_T;

#define QZ_MAX_STRING_LENGTH 64
typedef struct QzSoftwareVersionInfo_S {
    QzSoftwareComponentType_T component_type;
    unsigned char component_name[QZ_MAX_STRING_LENGTH];
    unsigned int major_version;
    unsigned int minor_version;
    unsigned int patch_version;
    unsigned int build_number;
    unsigned char reserved[52];
} QzSoftwareVersionInfo_T;

typedef struct QzCrc64Config_S {
    uint64_t polynomial;
    uint64_t initial_value;
    uint32_t reflect_in;
    uint32_t reflect_out;
    uint64_t xor_out;
} QzCrc64Config_T;

typedef struct QzCrc32Config_S {
    uint32_t polynomial;
    uint32_t initial_value;
    uint32_t reflect_in;
    uint32_t reflect_out;
    uint32_t xor_out;
} QzCrc32Config_T;

#define QZ_INPUT_CRC_VALID_BIT (4)
#define QZ_INPUT_CRC_VALID_MASK (1 << QZ_INPUT_CRC_VALID_BIT)
#define QZ_OUTPUT_CRC_VALID_BIT (8)
#define QZ_OUTPUT_CRC_VALID_MASK (1 << QZ_OUTPUT_CRC_VALID_BIT)
#define QZ_CRC32_VALID_BIT (12)
#define QZ_CRC32_VALID_MASK (1 << QZ_CRC32_VALID_BIT)
#define QZ_CRC64_VALID_BIT (16)
#define QZ_CRC64_VALID_MASK (1 << QZ_CRC64_VALID_BIT)
#define QZ_CRC32_VALID(f) ((f & QZ_CRC32_VALID_MASK) && !(f & QZ_CRC64_VALID_MASK))
#define QZ_CRC64_VALID(f) ((f & QZ_CRC64_VALID_MASK) && !(f & QZ_CRC32_VALID_MASK))
#define QZ_INPUT_CRC_VALID(f) (f & QZ_INPUT_CRC_VALID_MASK)
#define QZ_OUTPUT_CRC_VALID(f) (f & QZ_OUTPUT_CRC_VALID_BIT)

typedef struct QzCrcResult_S {
    int status;
    uint32_t valid_flags;
    union { uint32_t *crc_32; uint64_t *crc_64; } in_crc;
    union { uint32_t *crc_32; uint64_t *crc_64; } out_crc;
} QzCrcResult_T;

typedef struct QzResult_S {
    int status;
    void *cb_tag;
    unsigned int src_len;
    unsigned int dest_len;
    uint64_t ext_rc;
    QzCrcResult_T *crc;
    void *extension_result;
} QzResult_T;

typedef int (*qzAsyncCallbackFn)(QzResult_T *res);
typedef void *QzMetadataBlob_T;

typedef enum QzLogLevel_E {
    LOG_NONE = 0, LOG_FATAL, LOG_ERROR, LOG_WARNING, LOG_INFO, LOG_DEBUG1, LOG_DEBUG2, LOG_DEBUG3
} QzLogLevel_T;

/* ---- library / session life cycle ----
 * qzInit          reference src/qatzip.c:630     bring up the process-wide device context
 * qzSetupSession* reference src/qatzip.c:1118-1345 validate parameters, attach them to `sess`
 * qzTeardownSession :2673, qzClose :2748 */
QzLogLevel_T qzSetLogLevel(QzLogLevel_T level);
int qzInit(QzSession_T *sess, unsigned char sw_backup);
int qzSetupSession(QzSession_T *sess, QzSessionParams_T *params);
int qzSetupSessionDeflate(QzSession_T *sess, QzSessionParamsDeflate_T *params);
int qzSetupSessionLZ4(QzSession_T *sess, QzSessionParamsLZ4_T *params);
int qzSetupSessionLZ4S(QzSession_T *sess, QzSessionParamsLZ4S_T *params);
int qzSetupSessionDeflateExt(QzSession_T *sess, QzSessionParamsDeflateExt_T *params);
int qzTeardownSession(QzSession_T *sess);
int qzClose(QzSession_T *sess);
int qzGetStatus(QzSession_T *sess, QzStatus_T *status);
int qzGetDeflateEndOfStream(QzSession_T *sess, unsigned char *endofstream);

/* ---- one-shot compress: reference src/qatzip.c:1842-2097 ----
 * in:  *src_len = bytes available, *dest_len = capacity, last in {0,1}
 * out: *src_len = bytes consumed (whole chunks), *dest_len = bytes produced
 * `crc` (deflate formats) accumulates the CRC-32 of everything consumed, 0 restarts it. */
int qzCompress(QzSession_T *sess, const unsigned char *src, unsigned int *src_len,
               unsigned char *dest, unsigned int *dest_len, unsigned int last);
int qzCompressExt(QzSession_T *sess, const unsigned char *src, unsigned int *src_len,
                  unsigned char *dest, unsigned int *dest_len, unsigned int last, uint64_t *ext_rc);
int qzCompressCrc(QzSession_T *sess, const unsigned char *src, unsigned int *src_len,
                  unsigned char *dest, unsigned int *dest_len, unsigned int last, unsigned long *crc);
int qzCompressCrcExt(QzSession_T *sess, const unsigned char *src, unsigned int *src_len,
                     unsigned char *dest, unsigned int *dest_len, unsigned int last,
                     unsigned long *crc, uint64_t *ext_rc);
int qzCompressCrc64(QzSession_T *sess, const unsigned char *src, unsigned int *src_len,
                    unsigned char *dest, unsigned int *dest_len, unsigned int last, uint64_t *crc);
int qzCompressCrc64Ext(QzSession_T *sess, const unsigned char *src, unsigned int *src_len,
                       unsigned char *dest, unsigned int *dest_len, unsigned int last,
                       uint64_t *crc, uint64_t *ext_rc);
int qzCompressWithMetadataExt(QzSession_T *sess, const unsigned char *src, unsigned int *src_len,
                              unsigned char *dest, unsigned int *dest_len, unsigned int last,
                              uint64_t *ext_rc, QzMetadataBlob_T metadata,
                              uint32_t hw_buff_sz_override, uint32_t comp_thrshold);
int qzCompress2(QzSession_T *sess, const unsigned char *src, unsigned char *dest,
                qzAsyncCallbackFn callback, QzResult_T *qzResults);

/* ---- one-shot decompress: reference src/qatzip.c:2422-2671 ---- */
int qzDecompress(QzSession_T *sess, const unsigned char *src, unsigned int *src_len,
                 unsigned char *dest, unsigned int *dest_len);
int qzDecompressExt(QzSession_T *sess, const unsigned char *src, unsigned int *src_len,
                    unsigned char *dest, unsigned int *dest_len, uint64_t *ext_rc);
int qzDecompressCrc(QzSession_T *sess, const unsigned char *src, unsigned int *src_len,
                    unsigned char *dest, unsigned int *dest_len, unsigned long *crc);
int qzDecompressCrcExt(QzSession_T *sess, const unsigned char *src, unsigned int *src_len,
                       unsigned char *dest, unsigned int *dest_len, unsigned long *crc,
                       uint64_t *ext_rc);
int qzDecompressCrc64(QzSession_T *sess, const unsigned char *src, unsigned int *src_len,
                      unsigned char *dest, unsigned int *dest_len, uint64_t *crc);
int qzDecompressCrc64Ext(QzSession_T *sess, const unsigned char *src, unsigned int *src_len,
                         unsigned char *dest, unsigned int *dest_len, uint64_t *crc,
                         uint64_t *ext_rc);
int qzDecompressWithMetadataExt(QzSession_T *sess, const unsigned char *src, unsigned int *src_len,
                                unsigned char *dest, unsigned int *dest_len, uint64_t *ext_rc,
                                QzMetadataBlob_T metadata, uint32_t hw_buff_sz_override);
int qzDecompress2(QzSession_T *sess, const unsigned char *src, unsigned char *dest,
                  qzAsyncCallbackFn callback, QzResult_T *qzResults);

/* sizing: reference include/qatzip.h:2043-2044, src/qatzip.c:3022-3068 */
#define QZ_SKID_PAD_SZ 48
#define QZ_COMPRESSED_SZ_OF_EMPTY_FILE 34
unsigned int qzMaxCompressedLength(unsigned int src_sz, QzSession_T *sess);

/* process-wide defaults: reference src/qatzip.c:2780-2940 */
int qzSetDefaults(QzSessionParams_T *defaults);
int qzSetDefaultsDeflate(QzSessionParamsDeflate_T *defaults);
int qzSetDefaultsLZ4(QzSessionParamsLZ4_T *defaults);
int qzSetDefaultsLZ4S(QzSessionParamsLZ4S_T *defaults);
int qzSetDefaultsDeflateExt(QzSessionParamsDeflateExt_T *defaults);
int qzGetDefaults(QzSessionParams_T *defaults);
int qzGetDefaultsDeflate(QzSessionParamsDeflate_T *defaults);
int qzGetDefaultsLZ4(QzSessionParamsLZ4_T *defaults);
int qzGetDefaultsLZ4S(QzSessionParamsLZ4S_T *defaults);
int qzGetDefaultsDeflateExt(QzSessionParamsDeflateExt_T *defaults);

/* pinned memory: reference src/qatzip_mem.c:169 (qzMalloc), :226 (qzFree), :102 (qzMemFindAddr).
 * PINNED_MEM requests come from cudaMallocHost pages, which the engine DMAs without staging. */
void *qzMalloc(size_t sz, int numa, int force_pinned);
void qzFree(void *m);
int qzMemFindAddr(unsigned char *a);
int qzAllocateMetadata(QzMetadataBlob_T *metadata, size_t data_size, uint32_t hw_buff_sz);
int qzFreeMetadata(QzMetadataBlob_T metadata);

/* ---- stream API: reference include/qatzip.h:2358-2379, src/qatzip_stream.c:403,599,751 ---- */
typedef struct QzStream_S {
    unsigned int in_sz;          /* in: bytes offered at `in`; out: bytes not yet taken */
    unsigned int out_sz;         /* in: room at `out`; out: bytes written */
    unsigned char *in;
    unsigned char *out;
    unsigned int pending_in;     /* bytes buffered inside the stream, not yet coded */
    unsigned int pending_out;    /* coded bytes waiting for room at `out` */
    QzCrcType_T crc_type;
    unsigned int crc_32;         /* CRC-32 of all stream input so far */
    unsigned long long reserved;
    void *opaque;
} QzStream_T;
int qzCompressStream(QzSession_T *sess, QzStream_T *strm, unsigned int last);
int qzDecompressStream(QzSession_T *sess, QzStream_T *strm, unsigned int last);
int qzEndStream(QzSession_T *sess, QzStream_T *strm);

/* declared by the reference, implemented there as stubs returning QZ_FAIL
 * (reference src/qatzip.c:3070-3081) */
int qzGetSoftwareComponentVersionList(QzSoftwareVersionInfo_T *api_info, unsigned int *num_elem);
int qzGetSoftwareComponentCount(unsigned int *num_elem);

/* declared by the reference header but defined nowhere in its sources (SURVEY.md section 2
 * row 24); exported here as QZ_NOT_SUPPORTED stubs so that any caller still links */
int qzGetSessionCrc64Config(QzSession_T *sess, QzCrc64Config_T *crc64_config);
int qzGetSessionCrc32Config(QzSession_T *sess, QzCrc32Config_T *crc32_config);
int qzSetSessionCrc64Config(QzSession_T *sess, QzCrc64Config_T *crc64_config);
int qzSetSessionCrc32Config(QzSession_T *sess, QzCrc32Config_T *crc32_config);
int qzMetadataBlockRead(uint32_t block_num, QzMetadataBlob_T metadata, uint32_t *block_offset,
                        uint32_t *block_size, uint32_t *block_flags, uint32_t *block_hash);
int qzMetadataBlockWrite(uint32_t block_num, QzMetadataBlob_T metadata, uint32_t *block_offset,
                         uint32_t *block_size, uint32_t *block_flags, uint32_t *block_hash);
int qzMetadataBlockGetCrc64(uint32_t block_num, QzMetadataBlob_T metadata, uint64_t *input_crc,
                            uint64_t *output_crc);
int qzMetadataBlockGetCrc32(uint32_t block_num, QzMetadataBlob_T metadata, uint32_t *input_crc,
                            uint32_t *output_crc);

#ifdef __cplusplus
}
#endif
#endif
