/* warp_emu.h -- TEST INFRASTRUCTURE, never part of libqatzip.so.
 *
 * A small SIMT emulator that lets the kernel sources under qatzip_b200/csrc/ be compiled by g++ and
 * run on the CPU, so that their warp-level logic (ballots, shuffles, match.any, shared-memory hand-overs,
 * the piece-buffer pool) can be checked in the `-m "not gpu"` suite and while developing without a GPU.
 * It measures nothing and ships nowhere: the product has no CPU path.
 *
 * Model: one CTA at a time; every CUDA thread is a cooperative fiber with its own stack.  A fiber runs
 * until it reaches a warp collective (__shfl_sync, __ballot_sync, __match_any_sync, __reduce_or_sync,
 * __syncwarp), a CTA barrier or __nanosleep, then the next fiber runs.  Lanes of a warp therefore execute
 * the code between two collectives one after the other -- a legal schedule under independent thread
 * scheduling, and one that makes a missing __syncwarp visible as a wrong result rather than hiding it
 * behind lockstep execution.  A collective reached with different operations by the lanes of one warp,
 * or a state where no fiber can run, aborts with a message.  Dynamic shared memory is poisoned before
 * every CTA.  Global memory is ordinary host memory.
 */
#ifndef QZ_WARP_EMU_H
#define QZ_WARP_EMU_H
#define QZ_WARP_EMU 1
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdio.h>
#include <stdlib.h>
#include <functional>

/* ---- CUDA keywords ---- */
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) alignas(n)

struct EmuDim3 { unsigned x, y, z; };
extern EmuDim3 threadIdx, blockIdx, blockDim, gridDim;

struct alignas(16) uint4 { uint32_t x, y, z, w; };
struct alignas(8) uint2 { uint32_t x, y; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 v = { x, y, z, w }; return v; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { uint2 v = { x, y }; return v; }

/* ---- core (warp_emu.cpp) ---- */
namespace emu {
enum Op { OP_SYNCWARP = 1, OP_SHFL, OP_SHFL_UP, OP_SHFL_DOWN, OP_SHFL_XOR, OP_BALLOT, OP_MATCH, OP_REDUCE };
unsigned lane();
/* deposits v, waits for the live lanes of the warp, returns the 32 deposited values (dead lanes: stale) and the live mask */
const uint64_t *exchange(int op, uint64_t v, uint32_t *live);
void syncwarp();
void syncthreads();
void yield_sleep();
void named_barrier(unsigned id, unsigned count);
uint8_t *dyn_smem();
/* run body() once per thread of a grid x block launch with smem bytes of dynamic shared memory */
void launch(unsigned grid, unsigned block, size_t smem, const std::function<void()> &body);
unsigned long long collectives();     /* count since process start, for curiosity */
}

#define QZ_EMU_DYN_SMEM(name) uint8_t *name = emu::dyn_smem()

/* ---- warp collectives ---- */
template <class T> static inline uint64_t emu_pack(T v) { static_assert(sizeof(T) <= 8, "shuffle payload"); uint64_t u = 0; memcpy(&u, &v, sizeof(T)); return u; }
template <class T> static inline T emu_unpack(uint64_t u) { T v; memcpy(&v, &u, sizeof(T)); return v; }

template <class T> static inline T __shfl_sync(uint32_t, T v, int src, int = 32)
{
    uint32_t live; const uint64_t *x = emu::exchange(emu::OP_SHFL, emu_pack(v), &live);
    const unsigned s = (unsigned)src & 31u;
    return ((live >> s) & 1u) ? emu_unpack<T>(x[s]) : v;
}
template <class T> static inline T __shfl_up_sync(uint32_t, T v, unsigned d, int = 32)
{
    uint32_t live; const uint64_t *x = emu::exchange(emu::OP_SHFL_UP, emu_pack(v), &live);
    const unsigned l = emu::lane();
    return l >= d ? emu_unpack<T>(x[l - d]) : v;
}
template <class T> static inline T __shfl_down_sync(uint32_t, T v, unsigned d, int = 32)
{
    uint32_t live; const uint64_t *x = emu::exchange(emu::OP_SHFL_DOWN, emu_pack(v), &live);
    const unsigned l = emu::lane();
    return l + d < 32 ? emu_unpack<T>(x[l + d]) : v;
}
template <class T> static inline T __shfl_xor_sync(uint32_t, T v, int m, int = 32)
{
    uint32_t live; const uint64_t *x = emu::exchange(emu::OP_SHFL_XOR, emu_pack(v), &live);
    return emu_unpack<T>(x[(emu::lane() ^ (unsigned)m) & 31u]);
}
static inline uint32_t __ballot_sync(uint32_t, int pred)
{
    uint32_t live; const uint64_t *x = emu::exchange(emu::OP_BALLOT, pred ? 1u : 0u, &live);
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) if (((live >> i) & 1u) && x[i]) r |= 1u << i;
    return r;
}
static inline uint32_t __match_any_sync(uint32_t, uint32_t v)
{
    uint32_t live; const uint64_t *x = emu::exchange(emu::OP_MATCH, v, &live);
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) if (((live >> i) & 1u) && (uint32_t)x[i] == v) r |= 1u << i;
    return r;
}
static inline uint32_t __reduce_or_sync(uint32_t, uint32_t v)
{
    uint32_t live; const uint64_t *x = emu::exchange(emu::OP_REDUCE, v, &live);
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) if ((live >> i) & 1u) r |= (uint32_t)x[i];
    return r;
}
static inline uint32_t __reduce_xor_sync(uint32_t, uint32_t v)
{
    uint32_t live; const uint64_t *x = emu::exchange(emu::OP_REDUCE, v, &live);
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) if ((live >> i) & 1u) r ^= (uint32_t)x[i];
    return r;
}
static inline uint32_t __reduce_max_sync(uint32_t, uint32_t v)
{
    uint32_t live; const uint64_t *x = emu::exchange(emu::OP_REDUCE, v, &live);
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) if (((live >> i) & 1u) && (uint32_t)x[i] > r) r = (uint32_t)x[i];
    return r;
}
static inline void __syncwarp(uint32_t = 0xffffffffu) { emu::syncwarp(); }
static inline void __syncthreads() { emu::syncthreads(); }
static inline void __nanosleep(unsigned) { emu::yield_sleep(); }
static inline void __threadfence_block() {}
static inline void __threadfence() {}
static inline long long clock64() { return 0; }

/* ---- atomics: fibers never run concurrently ---- */
template <class T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicSub(T *p, T v) { T o = *p; *p = o - v; return o; }
template <class T> static inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
template <class T> static inline T atomicAnd(T *p, T v) { T o = *p; *p = o & v; return o; }
template <class T> static inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> static inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
template <class T> static inline T atomicCAS(T *p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }

/* ---- integer intrinsics ---- */
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline uint32_t __brev(uint32_t v)
{
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
    return __builtin_bswap32(v);
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) { s &= 31u; return s ? (lo >> s) | (hi << (32 - s)) : lo; }
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t s) { s &= 31u; return s ? (hi << s) | (lo >> (32 - s)) : hi; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T> static inline T __ldcg(const T *p) { return *p; }

/* CUDA's overloaded min/max */
static inline uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }
static inline uint32_t max(uint32_t a, uint32_t b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline uint32_t min(uint32_t a, int b) { return min(a, (uint32_t)b); }
static inline uint32_t min(int a, uint32_t b) { return min((uint32_t)a, b); }
static inline uint32_t max(uint32_t a, int b) { return max(a, (uint32_t)b); }
static inline uint32_t max(int a, uint32_t b) { return max((uint32_t)a, b); }
static inline uint64_t min(uint64_t a, uint64_t b) { return a < b ? a : b; }
static inline uint64_t max(uint64_t a, uint64_t b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }

#endif
