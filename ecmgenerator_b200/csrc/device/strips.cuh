// Multi-GPU strip decomposition: device side of the per-tick halo exchange and agent migration.
//
// The world is cut into vertical strips [lo, hi) along x, one per rank (one process per GPU).
// Every rank holds the per-slot arrays for ALL global slots; static components (radius, speed,
// path) are replicated at load time, `active[slot]` is 1 only on the rank that OWNS the agent.
// All inter-agent coupling of the reference tick is the read of the k = 5 neighbours' pre-tick
// position / velocity / radius (ORCA.cpp:342-344), so one exchange per tick suffices:
//
//   k_pack            owned agents near a strip border -> halo entries for that neighbour;
//                     agents that left the strip -> migrant entries for the new owner, the old
//                     owner keeps them as "self ghosts" for the coming tick
//   NCCL send/recv    one fixed-size message per direction (header with the two counts)
//   k_unpack_migrants received migrants become owned agents of this rank
//   (peer transport: k_pack stores into the neighbour's inbox over NVLink, k_exchange_p2p publishes the
//    counts + sequence number, waits for both neighbours and adopts the migrants - no NCCL call)
//   k_ghost_count / k_ghost_scatter   received halo entries + self ghosts join the neighbour grid
//                     snapshot as ghosts (visible as neighbours, never updated here)
//
// Exactness: the grid holds every agent with x in [lo - halo, hi + halo).  After the neighbour
// search an owned agent whose 5th distance reaches beyond that range raises ECMGPU_ST_HALO_MISS
// (counted); with zero misses the result equals the single-GPU result bit for bit.
#pragma once
#include "tick.cuh"

namespace ecm {

struct HaloEntry {  // 20 B
    int slot;
    float x, y, vx, vy;
};
struct MigrantEntry {  // 48 B: every mutable per-slot component
    int slot;
    float x, y, vx, vy;
    float ax, ay;    // attraction point (kept on IRM failure, Simulator.cpp:573-587)
    float pvx, pvy;  // preferred velocity
    float fx, fy;    // force
    unsigned flags;  // bit 0: replan pending
};
struct MsgHeader {  // 16 B
    int n_halo, n_migrants, pad0, pad1;
};

// Compact walk (ECMGPU_COMPACT=1, off by default): the slots this rank may own, so that the per-slot kernels of a tick
// (pack, cell count, scatter) walk the owned share instead of every slot of the global crowd.  No slot is listed twice
// (`in_list`), inactive entries are skipped; adopted migrants are appended, agents that left stay listed (inactive) until
// the next rebuild.  list == nullptr: off.
struct WalkView {
    int* list;               // [max_agents]
    int* n;                  // entries
    unsigned char* in_list;  // [max_agents]
};

struct StripView {
    int enabled;
    int rank, n_ranks;
    float lo, hi, halo;  // lo = -inf on rank 0, hi = +inf on the last rank
    int cap_halo, cap_migr, cap_self;
    // message = [MsgHeader][HaloEntry x cap_halo][MigrantEntry x cap_migr]; index 0 = left, 1 = right
    unsigned char* send[2];   // where the entries of the outgoing message are written: a local staging buffer (NCCL /
                              // in-process transports) or the neighbour's inbox itself, mapped over NVLink (peer transport)
    unsigned char* recv[2];
    MsgHeader* send_hdr[2];   // local header the pack kernel counts in (== send[d] unless the peer transport is on)
    HaloEntry* self_ghost;
    int* self_ghost_n;
    int* g_key;   // [2*cap_halo + cap_self] cell key of ghost g
    int* g_rank;
    WalkView walk;
    __device__ __forceinline__ MsgHeader* hdr(unsigned char* m) const { return (MsgHeader*)m; }
    __device__ __forceinline__ HaloEntry* halo_of(unsigned char* m) const { return (HaloEntry*)(m + sizeof(MsgHeader)); }
    __device__ __forceinline__ MigrantEntry* migr_of(unsigned char* m) const {
        return (MigrantEntry*)(m + sizeof(MsgHeader) + sizeof(HaloEntry) * (size_t)cap_halo);
    }
};

__host__ __device__ inline size_t strip_msg_bytes(int cap_halo, int cap_migr) {
    return sizeof(MsgHeader) + sizeof(HaloEntry) * (size_t)cap_halo + sizeof(MigrantEntry) * (size_t)cap_migr;
}

// Ownership from position: used once when strips are installed.
__global__ void __launch_bounds__(256) k_assign_owner(int n_slots, AgentArrays ag, StripView sv) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_slots || !ag.active[i]) return;
    float x = ag.pos[i].x;
    if (!(x >= sv.lo && x < sv.hi)) ag.active[i] = 0;
}

constexpr int kPackBlock = 1024;

// One slot of k_pack (`owned` false: a thread without an agent - it still takes part in the CTA-wide reservations).
__device__ __forceinline__ void pack_slot(bool owned, int i, AgentArrays& ag, const StripView& sv, unsigned long long* counters, int* s_warp) {
    float2 p = make_float2(0.0f, 0.0f), v = p;
    if (owned) p = ag.pos[i];
    const int dir = !owned ? -1 : (p.x < sv.lo ? 0 : (p.x >= sv.hi ? 1 : -1));
    if (dir >= 0) {  // left the strip (rare: per-entry atomics): hand over, keep as a ghost for the coming tick
        v = ag.vel[i];
        unsigned char* m = sv.send[dir];
        int e = atomicAdd(&sv.send_hdr[dir]->n_migrants, 1);
        if (e < sv.cap_migr) {
            MigrantEntry me;
            me.slot = i; me.x = p.x; me.y = p.y; me.vx = v.x; me.vy = v.y;
            const float2 a = ag.attraction[i], pv = ag.prefvel[i], f = ag.force[i];
            me.ax = a.x; me.ay = a.y; me.pvx = pv.x; me.pvy = pv.y; me.fx = f.x; me.fy = f.y;
            me.flags = ag.replan_pending[i] ? 1u : 0u;
            sv.migr_of(m)[e] = me;
            ag.active[i] = 0;
            int g = atomicAdd(sv.self_ghost_n, 1);
            if (g < sv.cap_self) {
                HaloEntry he; he.slot = i; he.x = p.x; he.y = p.y; he.vx = v.x; he.vy = v.y;
                sv.self_ghost[g] = he;
            } else atomicAdd(&counters[C_TOTAL_HALO_MISS], 1ull);
        } else {
            atomicAdd(&counters[C_TOTAL_HALO_MISS], 1ull);  // message full: the agent stays here this tick
        }
    }
    const bool stays = owned && dir < 0;
#pragma unroll
    for (int d = 0; d < 2; d++) {
        const bool has = d == 0 ? sv.rank > 0 : sv.rank < sv.n_ranks - 1;
        if (!has) continue;  // uniform over the grid
        const bool near = stays && (d == 0 ? p.x < sv.lo + sv.halo : p.x >= sv.hi - sv.halo);
        const int e = cta_reserve(near, &sv.send_hdr[d]->n_halo, s_warp);
        if (near) {
            if (e < sv.cap_halo) {
                v = ag.vel[i];
                HaloEntry he; he.slot = i; he.x = p.x; he.y = p.y; he.vx = v.x; he.vy = v.y;
                sv.halo_of(sv.send[d])[e] = he;
            } else atomicAdd(&counters[C_TOTAL_HALO_MISS], 1ull);
        }
    }
}

__global__ void __launch_bounds__(kPackBlock) k_pack(int n_slots, AgentArrays ag, StripView sv, unsigned long long* counters) {
    __shared__ int s_warp[33];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    pack_slot(i < n_slots && ag.active[i], i, ag, sv, counters, s_warp);
}

// ---- compact walk -----------------------------------------------------------------------------------------------
// (Re)builds the list from the active flags: after host-side changes of ownership (loads, spawns, writes of the
// active flags, new strip borders).  `*walk.n` and `in_list` are cleared by the caller.
__global__ void __launch_bounds__(kPackBlock) k_walk_rebuild(int n_slots, const unsigned char* __restrict__ active, WalkView walk) {
    __shared__ int s_warp[33];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool mine = i < n_slots && active[i];
    const int e = cta_reserve(mine, walk.n, s_warp);
    if (mine) {
        walk.list[e] = i;
        walk.in_list[i] = 1;
    }
}

// Fixed grid, every CTA makes the same number of trips (the CTA-wide reservations need all threads): the launch does
// not depend on the list length, so the tick still replays as a CUDA graph.
__global__ void __launch_bounds__(kPackBlock) k_pack_walk(AgentArrays ag, StripView sv, unsigned long long* counters) {
    __shared__ int s_warp[33];
    const int n = *sv.walk.n;
    const int stride = gridDim.x * blockDim.x;
    const int trips = (n + stride - 1) / stride;
    for (int it = 0; it < trips; it++) {
        const int idx = it * stride + blockIdx.x * blockDim.x + threadIdx.x;
        const int i = idx < n ? sv.walk.list[idx] : 0;
        pack_slot(idx < n && ag.active[i], i, ag, sv, counters, s_warp);
    }
}

// k_collect_owned (tick.cuh) over the list: the records of ecmgpu_update_io_owned.
__global__ void __launch_bounds__(kCollectBlock) k_collect_owned_walk(WalkView walk, const unsigned char* __restrict__ active, const float2* __restrict__ pos,
                                                                      const float2* __restrict__ vel, AgentRec* __restrict__ out, int out_cap, int* __restrict__ count,
                                                                      const int* __restrict__ ext_of) {
    __shared__ int s_warp[33];
    const int n = *walk.n;
    const int stride = gridDim.x * blockDim.x;
    const int trips = (n + stride - 1) / stride;
    for (int it = 0; it < trips; it++) {
        const int idx = it * stride + blockIdx.x * blockDim.x + threadIdx.x;
        const int i = idx < n ? walk.list[idx] : 0;
        const bool mine = idx < n && active[i];
        const int e = cta_reserve(mine, count, s_warp);
        if (mine && e < out_cap) {
            const float2 p = pos[i], v = vel[i];
            AgentRec r;
            r.slot = ext_of ? ext_of[i] : i; r.x = p.x; r.y = p.y; r.vx = v.x; r.vy = v.y;
            out[e] = r;
        }
    }
}

__device__ __forceinline__ void adopt_migrant(AgentArrays& ag, const MigrantEntry& me, const WalkView& walk) {
    if (walk.list && !walk.in_list[me.slot]) {  // one thread per migrant, one migrant per slot and tick: no race on the flag
        walk.in_list[me.slot] = 1;
        walk.list[atomicAdd(walk.n, 1)] = me.slot;
    }
    ag.pos[me.slot] = make_float2(me.x, me.y);
    ag.vel[me.slot] = make_float2(me.vx, me.vy);
    ag.attraction[me.slot] = make_float2(me.ax, me.ay);
    ag.prefvel[me.slot] = make_float2(me.pvx, me.pvy);
    ag.force[me.slot] = make_float2(me.fx, me.fy);
    ag.replan_pending[me.slot] = (me.flags & 1u) ? 1 : 0;
    ag.active[me.slot] = 1;
}

// Peer transport: the entries were stored straight into the neighbour's inbox by k_pack (NVLink
// peer stores).  This single-CTA kernel completes the exchange:
//   publish  counts, then - after a system-scope fence - the sequence number the neighbour spins on;
//   await    spin until both neighbours' messages carry this tick's sequence number;
//   adopt    received migrants become owned agents of this rank (k_unpack_migrants of the other transports).
// Inboxes are double-buffered by sequence parity; having seen the neighbour's message s-1 implies it
// has consumed our message s-2, so writing generation s & 1 never races with its reader.  The sequence
// number lives in device memory and is advanced here, so the tick can be replayed as a CUDA graph.
__global__ void __launch_bounds__(256) k_exchange_p2p(StripView sv, AgentArrays ag, int* seq_counter) {
    __shared__ int s_seq;
    if (threadIdx.x == 0) s_seq = *seq_counter;
    __syncthreads();
    const int seq = s_seq;
    const int d = threadIdx.x;
    const bool mine = d < 2 && (d == 0 ? sv.rank > 0 : sv.rank < sv.n_ranks - 1);
    if (mine) {
        MsgHeader* dst = (MsgHeader*)sv.send[d];
        const MsgHeader h = *sv.send_hdr[d];
        dst->n_halo = h.n_halo;
        dst->n_migrants = h.n_migrants;
        __threadfence_system();
        *(volatile int*)&dst->pad0 = seq;
    }
    __syncwarp();  // both messages are out before either lane starts to wait
    if (mine) {
        volatile int* flag = &((MsgHeader*)sv.recv[d])->pad0;
        while (*flag != seq) __nanosleep(100);
        __threadfence_system();
    }
    __syncthreads();
    if (threadIdx.x == 0) *seq_counter = seq + 1;
    for (int dd = 0; dd < 2; dd++) {
        const bool has = dd == 0 ? sv.rank > 0 : sv.rank < sv.n_ranks - 1;
        if (!has) continue;
        const unsigned char* m = sv.recv[dd];
        // the inbox was written by another GPU while this kernel ran: read it past the L1 (ld.cg)
        const int n = min(__ldcg(&((const MsgHeader*)m)->n_migrants), sv.cap_migr);
        const int* src = (const int*)(m + sizeof(MsgHeader) + sizeof(HaloEntry) * (size_t)sv.cap_halo);
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            union { MigrantEntry me; int w[sizeof(MigrantEntry) / 4]; } u;
#pragma unroll
            for (int k = 0; k < (int)(sizeof(MigrantEntry) / 4); k++) u.w[k] = __ldcg(src + (size_t)i * (sizeof(MigrantEntry) / 4) + k);
            adopt_migrant(ag, u.me, sv.walk);
        }
    }
}

__global__ void __launch_bounds__(256) k_unpack_migrants(AgentArrays ag, StripView sv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int d = 0; d < 2; d++) {
        const unsigned char* m = sv.recv[d];
        const int n = min(((const MsgHeader*)m)->n_migrants, sv.cap_migr);
        if (i < n) {
            const MigrantEntry me = ((const MigrantEntry*)(m + sizeof(MsgHeader) + sizeof(HaloEntry) * (size_t)sv.cap_halo))[i];
            adopt_migrant(ag, me, sv.walk);
        }
    }
}

// ghost index space: [0, cap_halo) from the left neighbour, [cap_halo, 2 cap_halo) from the right,
// [2 cap_halo, 2 cap_halo + cap_self) self ghosts
__device__ __forceinline__ bool ghost_entry(const StripView& sv, int g, HaloEntry& out) {
    if (g < 2 * sv.cap_halo) {
        const int d = g >= sv.cap_halo ? 1 : 0;
        const int k = g - d * sv.cap_halo;
        const unsigned char* m = sv.recv[d];
        if (k >= min(((const MsgHeader*)m)->n_halo, sv.cap_halo)) return false;
        out = ((const HaloEntry*)(m + sizeof(MsgHeader)))[k];
        return true;
    }
    const int k = g - 2 * sv.cap_halo;
    if (k >= min(*sv.self_ghost_n, sv.cap_self)) return false;
    out = sv.self_ghost[k];
    return true;
}

__device__ __forceinline__ void ghost_count_one(const StripView& sv, const GridParams& gp, int* __restrict__ cell_count, int g) {
    HaloEntry he;
    if (!ghost_entry(sv, g, he)) { sv.g_key[g] = -1; return; }
    float fx = (he.x - gp.x0) * gp.inv_cell, fy = (he.y - gp.y0) * gp.inv_cell;
    int cx = fx >= 0.0f ? (fx < (float)gp.w ? (int)fx : gp.w - 1) : 0;
    int cy = fy >= 0.0f ? (fy < (float)gp.h ? (int)fy : gp.h - 1) : 0;
    int k = cy * gp.w + cx;
    sv.g_key[g] = k;
    sv.g_rank[g] = atomicAdd(&cell_count[k], 1);
}

__device__ __forceinline__ void ghost_scatter_one(const StripView& sv, const int* __restrict__ cell_start, const AgentArrays& ag, const TickScratch& sc, int g) {
    const int k = sv.g_key[g];
    if (k < 0) return;
    HaloEntry he;
    ghost_entry(sv, g, he);
    const int p = cell_start[k] + sv.g_rank[g];
    sc.s_pos[p] = make_float2(he.x, he.y);
    sc.s_vel[p] = make_float2(he.vx, he.vy);
    sc.s_rad[p] = ag.radius[he.slot];
    sc.s_spd[p] = ag.speed[he.slot];
    sc.s_slot[p] = he.slot;
    sc.s_ghost[p] = 1;
}

__global__ void __launch_bounds__(256) k_ghost_count(StripView sv, GridParams gp, int* __restrict__ cell_count) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < 2 * sv.cap_halo + sv.cap_self) ghost_count_one(sv, gp, cell_count, g);
}

__global__ void __launch_bounds__(256) k_ghost_scatter(StripView sv, const int* __restrict__ cell_start, AgentArrays ag, TickScratch sc) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < 2 * sv.cap_halo + sv.cap_self) ghost_scatter_one(sv, cell_start, ag, sc, g);
}

// k_bin_count / k_scatter over the list (tick.cuh has the all-slots versions), with the ghosts in the same launch: at a
// rank's share of the crowd every launch saved is a few per cent of the tick.
__global__ void __launch_bounds__(256) k_bin_count_walk(StripView sv, const unsigned char* __restrict__ active, const float2* __restrict__ pos, GridParams gp,
                                                        int* __restrict__ cell_count, int* __restrict__ key, int* __restrict__ rank,
                                                        unsigned* __restrict__ status, unsigned long long* __restrict__ counters) {
    const int n = *sv.walk.n;
    const int stride = gridDim.x * blockDim.x, tid = blockIdx.x * blockDim.x + threadIdx.x;
    for (int idx = tid; idx < n; idx += stride) {
        const int i = sv.walk.list[idx];
        if (!active[i]) { key[i] = -1; continue; }
        const int k = grid_key(gp, pos[i], status, counters, i);
        key[i] = k;
        if (k >= 0) rank[i] = atomicAdd(&cell_count[k], 1);
    }
    const int ng = 2 * sv.cap_halo + sv.cap_self;
    for (int g = tid; g < ng; g += stride) ghost_count_one(sv, gp, cell_count, g);
}

__global__ void __launch_bounds__(256) k_scatter_walk(StripView sv, const int* __restrict__ key, const int* __restrict__ rank,
                                                      const int* __restrict__ cell_start, AgentArrays ag, TickScratch sc) {
    const int n = *sv.walk.n;
    const int stride = gridDim.x * blockDim.x, tid = blockIdx.x * blockDim.x + threadIdx.x;
    for (int idx = tid; idx < n; idx += stride) {
        const int i = sv.walk.list[idx];
        const int k = key[i];
        if (k < 0) continue;
        const int p = cell_start[k] + rank[i];
        sc.s_slot[p] = i;
        sc.s_pos[p] = ag.pos[i];
        sc.s_vel[p] = ag.vel[i];
        sc.s_rad[p] = ag.radius[i];
        sc.s_spd[p] = ag.speed[i];
        sc.s_ghost[p] = 0;
    }
    const int ng = 2 * sv.cap_halo + sv.cap_self;
    for (int g = tid; g < ng; g += stride) ghost_scatter_one(sv, cell_start, ag, sc, g);
}

}  // namespace ecm
