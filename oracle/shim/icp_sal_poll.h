/* oracle/shim -- TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_SHIM_ICP_SAL_POLL_H
#define ORACLE_SHIM_ICP_SAL_POLL_H
#include "cpa.h"
CpaStatus icp_sal_DcPollInstance(CpaInstanceHandle h, Cpa32U response_quota);
CpaStatus icp_sal_poll_device_events(void);
#endif
