/* warp_emu.cpp -- fiber scheduler behind warp_emu.h (test infrastructure; see that header). x86-64 only. */
#include "warp_emu.h"
#include <sys/mman.h>
#include <dlfcn.h>
#include <vector>

EmuDim3 threadIdx, blockIdx, blockDim, gridDim;

/* void emu_switch(void **save_sp, void *load_sp): callee-saved registers on the old stack, switch, restore */
extern "C" void emu_switch(void **save_sp, void *load_sp);
asm(".text\n.globl emu_switch\n.type emu_switch,@function\nemu_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n  ret\n"
    ".size emu_switch, .-emu_switch\n");

namespace emu {
namespace {
enum St { RUN, WAIT_WARP, WAIT_CTA, WAIT_NAMED, DONE };
struct Fiber { void *sp; unsigned tid; St st; unsigned gen; void *site; unsigned bar; bool slept; };
struct Warp { unsigned nlive, arrived; uint32_t live; int op; void *site; unsigned gen0; uint32_t snap[2]; uint64_t x[2][32]; };

const size_t STACK = 256 << 10;
const size_t SMEM_MAX = 256 << 10;
std::vector<Fiber> fibers;
std::vector<Warp> warps;
char *stacks; size_t stacks_sz;
alignas(128) uint8_t smem[SMEM_MAX];
void *sched_sp;
Fiber *cur;
unsigned cta_live, cta_arrived;
unsigned named_arrived[16];
bool rescan;
const std::function<void()> *body_fn;
unsigned long long ncoll;

[[noreturn]] void die(const char *what)
{
    fprintf(stderr, "warp_emu: %s (block %u thread %u)\n", what, blockIdx.x, cur ? cur->tid : 0u);
    abort();
}
void to_scheduler() { emu_switch(&cur->sp, sched_sp); }
void release_warp(unsigned w)
{
    warps[w].arrived = 0;
    warps[w].snap[warps[w].gen0 & 1] = warps[w].live;      /* who took part: a lane that exits later was still there */
    for (unsigned l = 0; l < 32; l++) { const unsigned t = w * 32 + l; if (t < fibers.size() && fibers[t].st == WAIT_WARP) fibers[t].st = RUN; }
}
void release_cta()
{
    cta_arrived = 0;
    for (auto &f : fibers) if (f.st == WAIT_CTA) f.st = RUN;
}
void warp_arrive(int op, void *site)
{
    Warp &w = warps[cur->tid >> 5];
    cur->site = site;
    if (w.arrived == 0) { w.op = op; w.site = site; w.gen0 = cur->gen; }
    else if (w.op != op || w.gen0 != cur->gen) {
        Dl_info a, b;                 /* offsets for addr2line -e libqzemu.so */
        const long oa = dladdr(w.site, &a) ? (long)((char *)w.site - (char *)a.dli_fbase) : 0, ob = dladdr(site, &b) ? (long)((char *)site - (char *)b.dli_fbase) : 0;
        fprintf(stderr, "warp_emu: op %d at +0x%lx (first arrival, its collective #%u) vs op %d at +0x%lx (lane %u, its collective #%u)\n", w.op, oa, w.gen0, op, ob, cur->tid & 31, cur->gen);
        for (unsigned l = 0; l < 32; l++) { const Fiber &f = fibers[(cur->tid & ~31u) + l]; Dl_info d; fprintf(stderr, "  lane %2u state %d collective #%u at +0x%lx\n", l, (int)f.st, f.gen, dladdr(f.site, &d) ? (long)((char *)f.site - (char *)d.dli_fbase) : 0L); }
        die("lanes of one warp reached different collectives (divergent warp-synchronous code)");
    }
    if (++w.arrived == w.nlive) {
        /* the lane that completes the rendezvous does not run on: the warp restarts from its lowest lane, so that the
         * order of the lanes through the next stretch of code never depends on who arrived last */
        release_warp(cur->tid >> 5);
        rescan = true;
        to_scheduler();
    } else { cur->st = WAIT_WARP; to_scheduler(); }
}
void fiber_main()
{
    (*body_fn)();
    /* thread exit: it no longer takes part in barriers */
    Warp &w = warps[cur->tid >> 5];
    w.nlive--; w.live &= ~(1u << (cur->tid & 31)); cta_live--;
    cur->st = DONE;
    if (w.nlive && w.arrived == w.nlive) release_warp(cur->tid >> 5);
    if (cta_live && cta_arrived == cta_live) release_cta();
    to_scheduler();
    die("finished fiber resumed");
}
}  // namespace

unsigned lane() { return cur->tid & 31; }
uint8_t *dyn_smem() { return smem; }
unsigned long long collectives() { return ncoll; }

const uint64_t *exchange(int op, uint64_t v, uint32_t *live)
{
    Warp &w = warps[cur->tid >> 5];
    uint64_t *buf = w.x[cur->gen & 1];
    buf[cur->tid & 31] = v;
    cur->gen++; ncoll++;
    warp_arrive(op, __builtin_return_address(0));
    *live = w.snap[cur->gen & 1];
    return buf;
}
void syncwarp() { cur->gen++; warp_arrive(OP_SYNCWARP, __builtin_return_address(0)); }
void syncthreads()
{
    if (++cta_arrived == cta_live) release_cta();
    else { cur->st = WAIT_CTA; to_scheduler(); }
}
void yield_sleep() { cur->slept = true; to_scheduler(); }
/* bar.sync id, count: `count` threads of the CTA meet at hardware barrier `id` (1..15) */
void named_barrier(unsigned id, unsigned count)
{
    if (id == 0 || id > 15 || count == 0 || count % 32) die("named barrier: bad id or thread count");
    if (++named_arrived[id] == count) {
        named_arrived[id] = 0;
        for (auto &f : fibers) if (f.st == WAIT_NAMED && f.bar == id) f.st = RUN;
    } else { cur->st = WAIT_NAMED; cur->bar = id; to_scheduler(); }
}

void launch(unsigned grid, unsigned block, size_t smem_bytes, const std::function<void()> &body)
{
    if (block == 0 || block > 1024 || smem_bytes > SMEM_MAX) { fprintf(stderr, "warp_emu: bad launch\n"); abort(); }
    if (stacks_sz < (size_t)block * STACK) {
        if (stacks) munmap(stacks, stacks_sz);
        stacks_sz = (size_t)block * STACK;
        stacks = (char *)mmap(NULL, stacks_sz, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (stacks == MAP_FAILED) { perror("warp_emu: mmap"); abort(); }
    }
    body_fn = &body;
    gridDim = { grid, 1, 1 }; blockDim = { block, 1, 1 };
    for (unsigned b = 0; b < grid; b++) {
        blockIdx = { b, 0, 0 };
        memset(smem, 0xCD, smem_bytes);                     /* dynamic shared memory starts undefined on the device */
        fibers.assign(block, Fiber());
        warps.assign((block + 31) / 32, Warp());
        cta_live = block; cta_arrived = 0;
        memset(named_arrived, 0, sizeof named_arrived);
        for (unsigned t = 0; t < block; t++) {
            Fiber &f = fibers[t];
            f.tid = t; f.st = RUN; f.gen = 0;
            Warp &w = warps[t >> 5]; w.nlive++; w.live |= 1u << (t & 31);
            /* initial frame: six callee-saved registers, then the entry point as return address; rsp = 8 mod 16 on entry */
            uintptr_t top = ((uintptr_t)(stacks + (size_t)(t + 1) * STACK)) & ~(uintptr_t)15;
            void **sp = (void **)top;
            *--sp = NULL;
            *--sp = (void *)&fiber_main;
            for (int i = 0; i < 6; i++) *--sp = NULL;
            f.sp = sp;
        }
        /* Warps take turns; inside a warp the runnable lanes always run in ascending lane order, and a warp keeps the
         * processor until none of its lanes can run.  The order in which the lanes of a warp pass through the code between
         * two collectives is therefore the same whatever the other warps do: results do not depend on the launch geometry
         * (the kernels' one intended race, equal-hash insertions inside a tile, is always won by the highest lane). */
        unsigned done = 0;
        const unsigned nw = (block + 31) / 32;
        while (done < block) {
            bool progress = false;
            for (unsigned w = 0; w < nw; w++) {
                for (bool ran = true; ran;) {
                    ran = false;
                    for (unsigned t = w * 32; t < block && t < w * 32 + 32; t++) {
                        Fiber &f = fibers[t];
                        if (f.st != RUN) continue;
                        cur = &f; threadIdx = { t, 0, 0 };
                        f.slept = false;
                        emu_switch(&sched_sp, f.sp);
                        if (f.st == DONE) done++;
                        if (!f.slept) ran = true;           /* a lane that only slept (spin-wait back-off) does not hold the warp */
                        progress = true;
                        if (rescan) { rescan = false; break; }      /* a rendezvous completed: back to the warp's lowest lane */
                    }
                }
            }
            if (!progress) { cur = NULL; die("deadlock: every live thread waits at a barrier that cannot complete"); }
        }
        cur = NULL;
    }
}
}  // namespace emu
