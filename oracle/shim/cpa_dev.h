/* oracle/shim/cpa_dev.h -- TEST INFRASTRUCTURE ONLY. The reference's internal header includes
 * only this file yet uses CpaDc* types, so pull both in here. */
#ifndef ORACLE_SHIM_CPA_DEV_H
#define ORACLE_SHIM_CPA_DEV_H
#include "cpa.h"
#include "cpa_dc.h"
#endif
