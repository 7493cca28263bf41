/* oracle/shim/cpa.h -- TEST INFRASTRUCTURE ONLY (never linked into libqatzip.so).
 * Minimal stand-in for Intel QAT's cpa.h so the reference's src/*.c compile
 * unmodified with no QAT stack; every device entry point fails, which makes
 * the reference take its software (zlib / liblz4) path -- see oracle/README.md. */
#ifndef ORACLE_SHIM_CPA_H
#define ORACLE_SHIM_CPA_H
#include <stdint.h>
#include <stddef.h>
typedef uint8_t Cpa8U; typedef uint16_t Cpa16U; typedef uint32_t Cpa32U; typedef uint64_t Cpa64U;
typedef int32_t Cpa32S; typedef int32_t CpaStatus;
typedef enum { CPA_FALSE = 0, CPA_TRUE = 1 } CpaBoolean;
#define CPA_STATUS_SUCCESS (0)
#define CPA_STATUS_FAIL (-1)
#define CPA_STATUS_RETRY (-2)
#define CPA_STATUS_RESOURCE (-3)
#define CPA_STATUS_INVALID_PARAM (-4)
#define CPA_STATUS_FATAL (-5)
#define CPA_STATUS_UNSUPPORTED (-6)
#define CPA_STATUS_RESTARTING (-7)
#define CPA_INSTANCE_HANDLE_SINGLE ((CpaInstanceHandle)0)
#define CPA_INST_ID_SIZE 128
#define CPA_INST_NAME_SIZE 64
#define CPA_INST_PART_NAME_SIZE 64
#define CPA_INST_SW_VERSION_SIZE 64
#define CPA_INST_VENDOR_NAME_SIZE 64
typedef void *CpaInstanceHandle;
typedef uint64_t CpaPhysicalAddr;
typedef CpaPhysicalAddr (*CpaVirtualToPhysical)(void *pVirtualAddr);
typedef struct _CpaFlatBuffer { Cpa32U dataLenInBytes; Cpa8U *pData; } CpaFlatBuffer;
typedef struct _CpaBufferList { Cpa32U numBuffers; CpaFlatBuffer *pBuffers; void *pUserData; void *pPrivateMetaData; } CpaBufferList;
typedef enum { CPA_ACC_SVC_TYPE_DATA_COMPRESSION = 3 } CpaAccelerationServiceType;
typedef enum { CPA_OPER_STATE_DOWN = 0, CPA_OPER_STATE_UP } CpaOperationalState;
typedef struct _CpaPhysicalInstanceId { Cpa16U packageId; Cpa16U acceleratorId; Cpa16U executionEngineId; Cpa16U busAddress; Cpa32U kptAcHandle; } CpaPhysicalInstanceId;
typedef struct _CpaInstanceInfo2 {
    CpaAccelerationServiceType accelerationServiceType;
    Cpa8U vendorName[CPA_INST_VENDOR_NAME_SIZE + 1];
    Cpa8U partName[CPA_INST_PART_NAME_SIZE + 1];
    Cpa8U swVersion[CPA_INST_SW_VERSION_SIZE + 1];
    Cpa8U instName[CPA_INST_NAME_SIZE + 1];
    Cpa8U instID[CPA_INST_ID_SIZE + 1];
    CpaPhysicalInstanceId physInstId;
    Cpa32U nodeAffinity;
    CpaOperationalState operState;
    CpaBoolean requiresPhysicallyContiguousMemory;
    CpaBoolean isPolled;
    CpaBoolean isOffloaded;
} CpaInstanceInfo2;
typedef enum { CPA_INSTANCE_EVENT_RESTARTING = 0, CPA_INSTANCE_EVENT_RESTARTED, CPA_INSTANCE_EVENT_FATAL_ERROR } CpaInstanceEvent;
#endif
