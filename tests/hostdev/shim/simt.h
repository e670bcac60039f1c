// TEST INFRASTRUCTURE: a SIMT emulator for the host build of the device code (-DHD_SIMT).
//
// The plain shim runs every kernel with ONE thread per block, which makes warp collectives identities and cannot
// exercise anything that needs co-operating lanes (the scan kernels, CTA-wide reservations, the warp-per-agent
// exhaustive neighbour search, the warp-synchronous loops with their real trip counts).  Here a CTA runs as `blockDim.x`
// FIBERS (ucontext) on one OS thread, round-robin, switching at barriers:
//   __syncthreads()                     barrier over the CTA's fibers that have not returned yet
//   __syncwarp / __shfl* / __ballot /   barrier over the warp's fibers (two per value exchange: publish, then read),
//   __any / __reduce_max                values travel through a per-warp slot array
// A fiber that returns leaves its barriers (CUDA counts exited threads as arrived).  If no fiber can run and not all
// have returned the launch aborts: a barrier some lanes never reach - undefined on the GPU - is caught here.
// Execution is deterministic (no OS threads), so plain loads and stores serve as atomics.  CTAs run one after another,
// which makes `static` a faithful `__shared__`.
#pragma once
#include <ucontext.h>

#include <cstdio>
#include <cstdlib>
#include <functional>
#include <vector>

namespace hdsimt {

struct Barrier {
    int expected = 0, arrived = 0;
    unsigned gen = 0;
};
struct Fiber {
    ucontext_t ctx;
    char* stack = nullptr;  // from the pool below: never zero-filled, so only the pages a kernel really uses are touched
    bool done = false;
    Barrier* waiting = nullptr;
    unsigned wait_gen = 0;
    int tid = 0;
};
struct Warp {
    Barrier bar;
    unsigned long long slot[32];
    unsigned char pred[32];
    unsigned alive = 0;
};
struct Cta {
    Barrier bar;
    std::vector<Warp> warps;
    std::vector<Fiber> fibers;
    ucontext_t sched;
    int current = -1;
    const std::function<void()>* body = nullptr;
};
inline thread_local Cta* g_cta = nullptr;  // per OS thread: the mock-runtime tests run one rank per thread
constexpr size_t kStack = 512 * 1024;  // the kernels keep a few KB of constraint arrays per thread
inline std::vector<char*>& stack_pool() {
    static thread_local std::vector<char*> pool;
    return pool;
}

inline void leave(Barrier& b) {
    b.expected--;
    if (b.expected > 0 && b.arrived >= b.expected) { b.arrived = 0; b.gen++; }
}
inline void trampoline() {
    Cta* c = g_cta;
    Fiber& f = c->fibers[c->current];
    (*c->body)();
    f.done = true;
    Warp& w = c->warps[f.tid >> 5];
    w.alive &= ~(1u << (f.tid & 31));
    w.pred[f.tid & 31] = 0;
    leave(w.bar);
    leave(c->bar);
    swapcontext(&f.ctx, &c->sched);
}
inline void wait(Barrier& b) {
    Cta* c = g_cta;
    if (++b.arrived >= b.expected) {  // the last one to arrive releases the others and goes on
        b.arrived = 0;
        b.gen++;
        return;
    }
    Fiber& f = c->fibers[c->current];
    f.waiting = &b;
    f.wait_gen = b.gen;
    swapcontext(&f.ctx, &c->sched);
}
inline Warp& my_warp() { return g_cta->warps[g_cta->fibers[g_cta->current].tid >> 5]; }
inline int my_lane() { return g_cta->fibers[g_cta->current].tid & 31; }

}  // namespace hdsimt

// launch geometry: blockIdx / threadIdx are rewritten by the scheduler whenever it resumes a fiber
struct HdDim3 { int x, y, z; };
static thread_local HdDim3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {1, 1, 1}, gridDim = {1, 1, 1};

// Runs `body` as grid x block threads.
inline void hd_simt_launch(int grid, int block, const std::function<void()>& body) {
    using namespace hdsimt;
    while ((int)stack_pool().size() < block) stack_pool().push_back((char*)malloc(kStack));
    gridDim.x = grid;
    blockDim.x = block;
    for (int b = 0; b < grid; b++) {
        Cta cta;
        cta.body = &body;
        cta.bar.expected = block;
        cta.warps.resize((block + 31) / 32);
        cta.fibers.resize(block);
        for (int t = 0; t < block; t++) {
            Warp& w = cta.warps[t >> 5];
            w.bar.expected++;
            w.alive |= 1u << (t & 31);
            w.pred[t & 31] = 0;
            Fiber& f = cta.fibers[t];
            f.tid = t;
            f.stack = stack_pool()[t];
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack;
            f.ctx.uc_stack.ss_size = kStack;
            f.ctx.uc_link = nullptr;
            makecontext(&f.ctx, (void (*)())trampoline, 0);
        }
        g_cta = &cta;
        int left = block;
        while (left > 0) {
            bool progressed = false;
            for (int t = 0; t < block; t++) {
                Fiber& f = cta.fibers[t];
                if (f.done) continue;
                if (f.waiting) {
                    if (f.waiting->gen == f.wait_gen) continue;  // its barrier has not opened yet
                    f.waiting = nullptr;
                }
                cta.current = t;
                blockIdx.x = b;
                threadIdx.x = t;
                swapcontext(&cta.sched, &f.ctx);
                progressed = true;
                if (f.done) left--;
            }
            if (!progressed) {
                fprintf(stderr, "hd_simt: deadlock in block %d: %d threads wait at barriers the others never reach\n", b, left);
                abort();
            }
        }
        g_cta = nullptr;
    }
    blockIdx.x = 0; threadIdx.x = 0; gridDim.x = 1; blockDim.x = 1;
}

// ---- the collectives ---------------------------------------------------------------------------------------
// After a switch the scheduler has rewritten threadIdx for another fiber: every collective restores it on return.
#define HD_RESTORE_TID() (threadIdx.x = hdsimt::g_cta->fibers[hdsimt::g_cta->current].tid)

static inline void __syncthreads() { hdsimt::wait(hdsimt::g_cta->bar); HD_RESTORE_TID(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { hdsimt::wait(hdsimt::my_warp().bar); HD_RESTORE_TID(); }
template <class T>
static inline T hd_exchange(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    hdsimt::Warp& w = hdsimt::my_warp();
    unsigned long long bits = 0;
    memcpy(&bits, &v, sizeof(T));
    w.slot[hdsimt::my_lane()] = bits;
    hdsimt::wait(w.bar);
    const unsigned long long got = w.slot[src_lane & 31];
    hdsimt::wait(w.bar);
    HD_RESTORE_TID();
    T r;
    memcpy(&r, &got, sizeof(T));
    return r;
}
template <class T>
static inline T __shfl_sync(unsigned, T v, int src) { return hd_exchange(v, src); }
template <class T>
static inline T __shfl_up_sync(unsigned, T v, int delta) {
    const int lane = hdsimt::my_lane();
    return hd_exchange(v, lane >= delta ? lane - delta : lane);  // lanes below delta keep their own value
}
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int mask) { return hd_exchange(v, hdsimt::my_lane() ^ mask); }
static inline unsigned __ballot_sync(unsigned, bool p) {
    hdsimt::Warp& w = hdsimt::my_warp();
    w.pred[hdsimt::my_lane()] = p ? 1 : 0;
    hdsimt::wait(w.bar);
    unsigned m = 0;
    for (int l = 0; l < 32; l++)
        if ((w.alive >> l & 1u) && w.pred[l]) m |= 1u << l;
    hdsimt::wait(w.bar);
    HD_RESTORE_TID();
    return m;
}
static inline bool __any_sync(unsigned mask, bool p) { return __ballot_sync(mask, p) != 0u; }
static inline int __reduce_max_sync(unsigned, int v) {
    hdsimt::Warp& w = hdsimt::my_warp();
    w.slot[hdsimt::my_lane()] = (unsigned long long)(long long)v;
    hdsimt::wait(w.bar);
    int m = v;
    for (int l = 0; l < 32; l++)
        if (w.alive >> l & 1u) m = std::max(m, (int)(long long)w.slot[l]);
    hdsimt::wait(w.bar);
    HD_RESTORE_TID();
    return m;
}
