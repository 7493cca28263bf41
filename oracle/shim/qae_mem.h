/* oracle/shim -- TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_SHIM_QAE_MEM_H
#define ORACLE_SHIM_QAE_MEM_H
#include <stddef.h>
#include <stdint.h>
void *qaeMemAllocNUMA(size_t size, int node, size_t phys_alignment_byte);
void qaeMemFreeNUMA(void **ptr);
uint64_t qaeVirtToPhysNUMA(void *pVirtAddr);
#endif
