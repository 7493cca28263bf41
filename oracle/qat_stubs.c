/* oracle/qat_stubs.c -- TEST INFRASTRUCTURE ONLY.
 * No-hardware stubs: icp_sal_userIsQatAvailable() says "no", so the reference's qzInit takes its
 * no-device branch (reference src/qatzip.c:694-699) and every qzCompress/qzDecompress runs the
 * reference's own software path (src/qatzip_sw.c: zlib / liblz4).  Nothing here is product code. */
#include <stdlib.h>
#include "cpa.h"
#include "cpa_dc.h"
#include "icp_sal_poll.h"
#include "icp_sal_user.h"
#include "qae_mem.h"
#include "numa.h"
#define F return CPA_STATUS_FAIL
CpaStatus cpaDcGetNumInstances(Cpa16U *n) { if (n) *n = 0; F; }
CpaStatus cpaDcGetInstances(Cpa16U n, CpaInstanceHandle *h) { (void)n; (void)h; F; }
CpaStatus cpaDcInstanceGetInfo2(const CpaInstanceHandle h, CpaInstanceInfo2 *i) { (void)h; (void)i; F; }
CpaStatus cpaDcQueryCapabilities(CpaInstanceHandle h, CpaDcInstanceCapabilities *c) { (void)h; (void)c; F; }
CpaStatus cpaDcInstanceSetNotificationCb(const CpaInstanceHandle h, const CpaDcInstanceNotificationCbFunc cb, void *t) { (void)h; (void)cb; (void)t; F; }
CpaStatus cpaDcBufferListGetMetaSize(const CpaInstanceHandle h, Cpa32U n, Cpa32U *s) { (void)h; (void)n; (void)s; F; }
CpaStatus cpaDcGetNumIntermediateBuffers(CpaInstanceHandle h, Cpa16U *n) { (void)h; (void)n; F; }
CpaStatus cpaDcSetAddressTranslation(const CpaInstanceHandle h, CpaVirtualToPhysical f) { (void)h; (void)f; F; }
CpaStatus cpaDcStartInstance(CpaInstanceHandle h, Cpa16U n, CpaBufferList **b) { (void)h; (void)n; (void)b; F; }
CpaStatus cpaDcStopInstance(CpaInstanceHandle h) { (void)h; F; }
CpaStatus cpaDcGetSessionSize(CpaInstanceHandle h, CpaDcSessionSetupData *sd, Cpa32U *a, Cpa32U *b) { (void)h; (void)sd; (void)a; (void)b; F; }
CpaStatus cpaDcInitSession(CpaInstanceHandle h, CpaDcSessionHandle s, CpaDcSessionSetupData *sd, CpaBufferList *c, CpaDcCallbackFn cb) { (void)h; (void)s; (void)sd; (void)c; (void)cb; F; }
CpaStatus cpaDcRemoveSession(const CpaInstanceHandle h, CpaDcSessionHandle s) { (void)h; (void)s; F; }
CpaStatus cpaDcCompressData2(CpaInstanceHandle h, CpaDcSessionHandle s, CpaBufferList *a, CpaBufferList *b, CpaDcOpData *o, CpaDcRqResults *r, void *t) { (void)h; (void)s; (void)a; (void)b; (void)o; (void)r; (void)t; F; }
CpaStatus cpaDcDecompressData(CpaInstanceHandle h, CpaDcSessionHandle s, CpaBufferList *a, CpaBufferList *b, CpaDcRqResults *r, CpaDcFlush f, void *t) { (void)h; (void)s; (void)a; (void)b; (void)r; (void)f; (void)t; F; }
CpaStatus cpaDcDeflateCompressBound(const CpaInstanceHandle h, CpaDcHuffType t, Cpa32U in, Cpa32U *out) { (void)h; (void)t; (void)in; (void)out; F; }
CpaStatus cpaDcLZ4CompressBound(const CpaInstanceHandle h, Cpa32U in, Cpa32U *out) { (void)h; (void)in; (void)out; F; }
CpaStatus cpaDcLZ4SCompressBound(const CpaInstanceHandle h, Cpa32U in, Cpa32U *out) { (void)h; (void)in; (void)out; F; }
CpaStatus icp_sal_DcPollInstance(CpaInstanceHandle h, Cpa32U q) { (void)h; (void)q; F; }
CpaStatus icp_sal_poll_device_events(void) { F; }
CpaStatus icp_sal_userStartMultiProcess(const char *n, CpaBoolean l) { (void)n; (void)l; F; }
CpaStatus icp_sal_userStop(void) { return CPA_STATUS_SUCCESS; }
CpaBoolean icp_sal_userIsQatAvailable(void) { return CPA_FALSE; }
CpaStatus icp_adf_get_numDevices(Cpa32U *n) { if (n) *n = 0; F; }
void *qaeMemAllocNUMA(size_t size, int node, size_t align) { (void)size; (void)node; (void)align; return NULL; }
void qaeMemFreeNUMA(void **p) { (void)p; }
uint64_t qaeVirtToPhysNUMA(void *p) { (void)p; return 0; }
int numa_node_of_cpu(int cpu) { (void)cpu; return 0; }
