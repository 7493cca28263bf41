/* oracle/shim -- TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_SHIM_ICP_SAL_USER_H
#define ORACLE_SHIM_ICP_SAL_USER_H
#include "cpa.h"
CpaStatus icp_sal_userStartMultiProcess(const char *pProcessName, CpaBoolean limitDevAccess);
CpaStatus icp_sal_userStop(void);
CpaBoolean icp_sal_userIsQatAvailable(void);
CpaStatus icp_adf_get_numDevices(Cpa32U *num_devices);
#endif
