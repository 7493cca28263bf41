/* oracle/shim -- TEST INFRASTRUCTURE ONLY (libnuma is absent from the image). */
#ifndef ORACLE_SHIM_NUMA_H
#define ORACLE_SHIM_NUMA_H
int numa_node_of_cpu(int cpu);
#endif
