/* oracle/shim/cpa_dc.h -- TEST INFRASTRUCTURE ONLY. Declarations the reference touches
 * from QAT's data-compression API; stubs in oracle/qat_stubs.c return failure. */
#ifndef ORACLE_SHIM_CPA_DC_H
#define ORACLE_SHIM_CPA_DC_H
#include "cpa.h"
#define CPA_DC_API_VERSION_NUM_MAJOR (3)
#define CPA_DC_API_VERSION_NUM_MINOR (2)
#define CPA_DC_API_VERSION_AT_LEAST(major, minor) \
    (CPA_DC_API_VERSION_NUM_MAJOR > major || (CPA_DC_API_VERSION_NUM_MAJOR == major && CPA_DC_API_VERSION_NUM_MINOR >= minor))
typedef void *CpaDcSessionHandle;
typedef enum { CPA_DC_L1 = 1, CPA_DC_L2, CPA_DC_L3, CPA_DC_L4, CPA_DC_L5, CPA_DC_L6, CPA_DC_L7, CPA_DC_L8, CPA_DC_L9, CPA_DC_L10, CPA_DC_L11, CPA_DC_L12 } CpaDcCompLvl;
typedef enum { CPA_DC_HT_STATIC = 0, CPA_DC_HT_PRECOMP, CPA_DC_HT_FULL_DYNAMIC } CpaDcHuffType;
typedef enum { CPA_DC_LZS = 0, CPA_DC_ELZS, CPA_DC_LZSS, CPA_DC_DEFLATE, CPA_DC_LZ4, CPA_DC_LZ4S } CpaDcCompType;
typedef enum { CPA_DC_NONE = 0, CPA_DC_CRC32, CPA_DC_ADLER32, CPA_DC_CRC32_ADLER32, CPA_DC_XXHASH32 } CpaDcChecksum;
typedef enum { CPA_DC_DIR_COMPRESS = 0, CPA_DC_DIR_DECOMPRESS, CPA_DC_DIR_COMBINED } CpaDcSessionDir;
typedef enum { CPA_DC_STATEFUL = 0, CPA_DC_STATELESS } CpaDcSessionState;
typedef enum { CPA_DC_ASB_DISABLED = 0, CPA_DC_ASB_STATIC_DYNAMIC, CPA_DC_ASB_UNCOMP_STATIC_DYNAMIC_WITH_STORED_HDRS, CPA_DC_ASB_UNCOMP_STATIC_DYNAMIC_WITH_NO_HDRS, CPA_DC_ASB_ENABLED } CpaDcAutoSelectBest;
typedef enum { CPA_DC_FLUSH_NONE = 0, CPA_DC_FLUSH_FINAL, CPA_DC_FLUSH_SYNC, CPA_DC_FLUSH_FULL } CpaDcFlush;
typedef enum { CPA_DC_SKIP_DISABLED = 0, CPA_DC_SKIP_AT_START, CPA_DC_SKIP_AT_END, CPA_DC_SKIP_STRIDE } CpaDcSkipMode;
typedef enum { CPA_DC_LZ4_MAX_BLOCK_SIZE_64K = 0, CPA_DC_LZ4_MAX_BLOCK_SIZE_256K, CPA_DC_LZ4_MAX_BLOCK_SIZE_1M, CPA_DC_LZ4_MAX_BLOCK_SIZE_4M } CpaDcCompLZ4BlockMaxSize;
typedef enum { CPA_DC_MIN_3_BYTE_MATCH = 0, CPA_DC_MIN_4_BYTE_MATCH } CpaDcCompMinMatch;
typedef enum {
    CPA_DC_OK = 0, CPA_DC_INVALID_BLOCK_TYPE = -1, CPA_DC_BAD_STORED_BLOCK_LEN = -2, CPA_DC_TOO_MANY_CODES = -3,
    CPA_DC_INCOMPLETE_CODE_LENS = -4, CPA_DC_REPEATED_LENS = -5, CPA_DC_MORE_REPEAT = -6, CPA_DC_BAD_LITLEN_CODES = -7,
    CPA_DC_BAD_DIST_CODES = -8, CPA_DC_INVALID_CODE = -9, CPA_DC_INVALID_DIST = -10, CPA_DC_OVERFLOW = -11,
    CPA_DC_SOFTERR = -12, CPA_DC_FATALERR = -13, CPA_DC_MAX_RESUBITERR = -14, CPA_DC_INCOMPLETE_FILE_ERR = -15,
    CPA_DC_WDOG_TIMER_ERR = -16, CPA_DC_EP_HARDWARE_ERR = -17, CPA_DC_VERIFY_ERROR = -18, CPA_DC_EMPTY_DYM_BLK = -19,
    CPA_DC_CRC_INTEG_ERR = -20, CPA_DC_LZ4_MAX_BLOCK_SIZE_EXCEEDED = -93, CPA_DC_LZ4_BLOCK_OVERFLOW_ERR = -95
} CpaDcReqStatus;
typedef struct _CpaDcSessionSetupData {
    CpaDcCompLvl compLevel; CpaDcCompType compType; CpaDcHuffType huffType; CpaDcAutoSelectBest autoSelectBestHuffmanTree;
    CpaDcSessionDir sessDirection; CpaDcSessionState sessState; Cpa32U windowSize; CpaDcCompMinMatch minMatch;
    CpaDcCompLZ4BlockMaxSize lz4BlockMaxSize; CpaBoolean lz4BlockChecksum; CpaBoolean lz4BlockIndependence;
    CpaDcChecksum checksum; CpaBoolean accumulateXXHash;
} CpaDcSessionSetupData;
typedef struct _CpaDcRqResults { CpaDcReqStatus status; Cpa32U produced; Cpa32U consumed; Cpa32U checksum; CpaBoolean endOfLastBlock; CpaBoolean dataUncompressed; } CpaDcRqResults;
typedef struct _CpaDcSkipData { CpaDcSkipMode skipMode; Cpa32U skipLength; Cpa32U strideLength; Cpa32U firstSkipOffset; } CpaDcSkipData;
typedef struct _CpaDcOpData { CpaDcFlush flushFlag; CpaBoolean compressAndVerify; CpaBoolean compressAndVerifyAndRecover; CpaBoolean integrityCrcCheck; CpaBoolean verifyHwIntegrityCrcs; CpaDcSkipData inputSkipData; CpaDcSkipData outputSkipData; void *pCrcData; } CpaDcOpData;
typedef struct _CpaDcInstanceCapabilities {
    CpaBoolean statefulLZSCompression, statefulLZSDecompression, statelessLZSCompression, statelessLZSDecompression;
    CpaBoolean statefulLZSSCompression, statefulLZSSDecompression, statelessLZSSCompression, statelessLZSSDecompression;
    CpaBoolean statefulELZSCompression, statefulELZSDecompression, statelessELZSCompression, statelessELZSDecompression;
    CpaBoolean statefulDeflateCompression, statefulDeflateDecompression, statelessDeflateCompression, statelessDeflateDecompression;
    CpaBoolean statelessLZ4Compression, statelessLZ4Decompression, statefulLZ4Decompression, statelessLZ4SCompression;
    CpaBoolean checksumCRC32, checksumAdler32, checksumXXHash32, dynamicHuffman, dynamicHuffmanBufferReq, precompiledHuffman;
    CpaBoolean autoSelectBestHuffmanTree; Cpa8U validWindowSizeMaskCompression, validWindowSizeMaskDecompression;
    Cpa32U internalHuffmanMem; CpaBoolean endOfLastBlock, reportParityError, batchAndPack, compressAndVerify, compressAndVerifyStrict,
    compressAndVerifyAndRecover, integrityCrcs;
} CpaDcInstanceCapabilities;
typedef void (*CpaDcCallbackFn)(void *callbackTag, CpaStatus status);
typedef void (*CpaDcInstanceNotificationCbFunc)(const CpaInstanceHandle instanceHandle, void *pCallbackTag, const CpaInstanceEvent instanceEvent);
CpaStatus cpaDcGetNumInstances(Cpa16U *pNumInstances);
CpaStatus cpaDcGetInstances(Cpa16U numInstances, CpaInstanceHandle *dcInstances);
CpaStatus cpaDcInstanceGetInfo2(const CpaInstanceHandle h, CpaInstanceInfo2 *info);
CpaStatus cpaDcQueryCapabilities(CpaInstanceHandle h, CpaDcInstanceCapabilities *caps);
CpaStatus cpaDcInstanceSetNotificationCb(const CpaInstanceHandle h, const CpaDcInstanceNotificationCbFunc cb, void *tag);
CpaStatus cpaDcBufferListGetMetaSize(const CpaInstanceHandle h, Cpa32U numBuffers, Cpa32U *pSizeInBytes);
CpaStatus cpaDcGetNumIntermediateBuffers(CpaInstanceHandle h, Cpa16U *pNumBuffers);
CpaStatus cpaDcSetAddressTranslation(const CpaInstanceHandle h, CpaVirtualToPhysical virtual2Physical);
CpaStatus cpaDcStartInstance(CpaInstanceHandle h, Cpa16U numBuffers, CpaBufferList **pIntermediateBuffers);
CpaStatus cpaDcStopInstance(CpaInstanceHandle h);
CpaStatus cpaDcGetSessionSize(CpaInstanceHandle h, CpaDcSessionSetupData *sd, Cpa32U *pSessionSize, Cpa32U *pContextSize);
CpaStatus cpaDcInitSession(CpaInstanceHandle h, CpaDcSessionHandle s, CpaDcSessionSetupData *sd, CpaBufferList *ctx, CpaDcCallbackFn cb);
CpaStatus cpaDcRemoveSession(const CpaInstanceHandle h, CpaDcSessionHandle s);
CpaStatus cpaDcCompressData2(CpaInstanceHandle h, CpaDcSessionHandle s, CpaBufferList *src, CpaBufferList *dst, CpaDcOpData *op, CpaDcRqResults *res, void *tag);
CpaStatus cpaDcDecompressData(CpaInstanceHandle h, CpaDcSessionHandle s, CpaBufferList *src, CpaBufferList *dst, CpaDcRqResults *res, CpaDcFlush flush, void *tag);
CpaStatus cpaDcDeflateCompressBound(const CpaInstanceHandle h, CpaDcHuffType huffType, Cpa32U inputSize, Cpa32U *outputSize);
CpaStatus cpaDcLZ4CompressBound(const CpaInstanceHandle h, Cpa32U inputSize, Cpa32U *outputSize);
CpaStatus cpaDcLZ4SCompressBound(const CpaInstanceHandle h, Cpa32U inputSize, Cpa32U *outputSize);
#endif
