/* qz_warp.cuh -- the few PTX-level helpers the kernels share: lane masks, L2 residency hints, the
 * dynamic shared-memory declaration.  Under QZ_WARP_EMU (tests/emu: the kernel sources compiled by g++
 * against a SIMT emulator -- test infrastructure only, never linked into libqatzip.so) they reduce to
 * plain loads and stores. */
#ifndef QZ_WARP_CUH
#define QZ_WARP_CUH
#include <stdint.h>

#ifdef QZ_WARP_EMU
#define QZ_DYN_SMEM(name) QZ_EMU_DYN_SMEM(name)
static inline uint32_t qz_lanemask_lt() { return (1u << emu::lane()) - 1u; }
static inline uint64_t l2_policy_keep() { return 0; }
static inline uint64_t l2_policy_stream() { return 0; }
static inline uint32_t tok_ld(const uint32_t *a, uint64_t) { return *a; }
static inline uint4 tok_ld4(const uint32_t *a, uint64_t) { return *reinterpret_cast<const uint4 *>(a); }
static inline void tok_st(uint32_t *a, uint32_t v, uint64_t) { *a = v; }
static inline void tok16_st(uint16_t *a, uint16_t v, uint64_t) { *a = v; }
static inline uint4 stream_ld16(const uint4 *a, uint64_t) { return *a; }
static inline void slot_st(uint32_t *a, uint32_t v) { *a = v; }
/* bulk copy + mbarrier: the copy is synchronous here, the barrier a count of completed phases */
static inline void qz_mbar_init(uint64_t *mb) { *mb = 0; }
static inline void qz_mbar_wait(uint64_t *mb, uint32_t k) { while (*reinterpret_cast<volatile uint64_t *>(mb) <= k) __nanosleep(0); }
static inline void qz_bulk_load_arrive(void *dst, const void *src, uint32_t bytes, uint64_t *mb) { memcpy(dst, src, bytes); *mb += 1; }
#else
/* whole words of compressed output: written once, read once by the framing kernel much later */
__device__ __forceinline__ void slot_st(uint32_t *a, uint32_t v)
{
#ifdef QZ_SLOT_STREAM
    asm volatile("st.global.cs.u32 [%0], %1;" :: "l"(a), "r"(v) : "memory");
#else
    *a = v;
#endif
}
#define QZ_DYN_SMEM(name) extern __shared__ __align__(16) uint8_t name[]
/* ---- TMA bulk copy global -> shared, completion on an mbarrier (one arrival + the copy's bytes per phase) ----
 * qz_mbar_init: one thread, before a CTA barrier.  qz_bulk_load_arrive: one thread; `bytes` a multiple of 16 (0: arrival
 * only), both addresses 16-byte aligned; everything that thread (and, after a __syncwarp, its warp) wrote before is visible
 * to whoever passes qz_mbar_wait for that phase.  qz_mbar_wait(mb, k): blocks until phase k (0, 1, 2 ...) is complete. */
__device__ __forceinline__ void qz_mbar_init(uint64_t *mb)
{
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(mb);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(a) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void qz_mbar_wait(uint64_t *mb, uint32_t k)
{
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(mb), parity = k & 1u;
    uint32_t ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void qz_bulk_load_arrive(void *dst, const void *src, uint32_t bytes, uint64_t *mb)
{
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(mb), d = (uint32_t)__cvta_generic_to_shared(dst);
    if (bytes) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(a), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(d), "l"(src), "r"(bytes), "r"(a) : "memory");
    } else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(a) : "memory");
}
__device__ __forceinline__ uint32_t qz_lanemask_lt() { uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

/* L2 residency control.  The token scratch is written once and read twice by the same SM within
 * microseconds, while the input streams through exactly once: tokens ask L2 to keep them
 * (evict_last), input lines are marked evict_first and skip L1, so the stream does not push the
 * tokens out to HBM.  .cg keeps token accesses coherent at L2 between lanes. */
__device__ __forceinline__ uint64_t l2_policy_keep() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t l2_policy_stream() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint32_t tok_ld(const uint32_t *a, uint64_t pol)
{
    uint32_t v; asm volatile("ld.global.cg.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(pol) : "memory"); return v;
}
__device__ __forceinline__ uint4 tok_ld4(const uint32_t *a, uint64_t pol)
{
    uint4 v;
    asm volatile("ld.global.cg.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(a), "l"(pol) : "memory");
    return v;
}
__device__ __forceinline__ void tok_st(uint32_t *a, uint32_t v, uint64_t pol)
{
    asm volatile("st.global.cg.L2::cache_hint.u32 [%0], %1, %2;" :: "l"(a), "r"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void tok16_st(uint16_t *a, uint16_t v, uint64_t pol)
{
    asm volatile("st.global.cg.L2::cache_hint.u16 [%0], %1, %2;" :: "l"(a), "h"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ uint4 stream_ld16(const uint4 *a, uint64_t pol)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(a), "l"(pol));
    return v;
}
#endif
#endif
