// Kernels of one simulation tick (Simulator::Update, /root/reference/ECMAgentSimulator/Simulator.cpp:314-323).
//
//   k_bin_count   cell key + rank of every active agent            (replaces KDTree::Construct, KDTree.cpp:22-57)
//   k_scan_onepass exclusive scan of the per-cell counts
//   k_scatter     counting-sort scatter: SoA snapshot of the pre-tick state in cell order
//   k_attract     arrival test, ECM point location, IRM attraction point, preferred velocity
//                 (UpdateAttractionPointSystem + ApplySteeringForce, Simulator.cpp:538-590, 638-657)
//   k_orca        exact 5-NN, obstacle gather, ORCA half-planes, LP / LP3D, velocity and position
//                 integration (ApplyObstacleAvoidanceForce, UpdateVelocitySystem, UpdatePositionSystem,
//                 Simulator.cpp:659-686, 619-635, 592-606)
//   k_fallback    the stragglers: warp-per-agent exhaustive neighbour search for agents whose ring budget ran
//                 out, and RandomizedLP3D + integration for the agents k_orca parked (orca.cuh Lp3dQueue)
//
// The snapshot makes the tick Jacobi-style exactly like the reference: every agent reads its
// neighbours' PRE-tick position / velocity / radius (ORCA.cpp:342-344) while new values are written
// to the slot arrays.  Agents destroyed on arrival this tick stay in the snapshot, i.e. remain
// visible as neighbours for this tick (KD-tree membership is frozen before arrivals, Simulator.cpp:319).
#pragma once
#include "locate.cuh"
#include "orca.cuh"

namespace ecm {

enum Counter {
    C_REPLAN_N = 0,     // entries in ev_replan since the last poll
    C_DESTROYED_N = 1,  // entries in ev_destroyed since the last poll
    C_FALLBACK_N = 2,   // entries in fb_list this tick (reset every tick)
    C_LP3D_N = 3,       // entries in the LP3D queue this tick (reset every tick, together with C_FALLBACK_N)
    C_TOTAL_FALLBACK = 4,
    C_TOTAL_OBST_OVF = 5,
    C_TOTAL_LP3D = 6,
    C_TOTAL_LOCFAIL = 7,
    C_TOTAL_REPLAN = 8,
    C_TOTAL_HALO_MISS = 9,
    C_TOTAL_KD_TIES = 10,  // kdtree.cuh: tree segments (> 16 elements) whose median tied on the split axis
    C_TOTAL_KD_SMALL_TIES = 11,  // the same in segments of up to 16 elements (resolved like libstdc++)
    C_TOTAL_EV_OVERFLOW = 12,
    C_TOTAL_NONFINITE = 13,      // agent-ticks skipped because the agent's position is not finite (see grid_key)    // events dropped because a queue was full (the host did not poll for max_agents events)
    C_COUNT = 16
};

// Path header of a slot: `off` is a multiple of 8 points (blocks of 8 segments share bbox index off/8 + b);
// the goal (last point, Simulator.cpp:554) is duplicated here so the arrival test touches no polyline data.
struct PathHdr {
    int off, len;
    float gx, gy;
};

// Mutable per-slot agent components (structure of arrays, indexed by slot = global agent id).
struct AgentArrays {
    float2* pos;          // Simulator::m_Positions
    float2* vel;          // m_Velocities
    float2* prefvel;      // m_PreferredVelocities
    float2* attraction;   // m_AttractionPoints
    float2* force;        // m_Forces
    float* radius;        // m_Clearances
    float* speed;         // m_PreferredSpeed
    unsigned char* active;        // m_ActiveAgents
    unsigned char* replan_pending;
    unsigned* status;
    int* cell;            // located ECM cell of the last tick
    int* nbr;             // [5*slot] neighbour slot ids of the last tick (optional)
    int* nbr_cnt;
    const PathHdr* path_hdr;  // (offset, length, goal) into path_pool
    const float2* path_pool;
    const float4* path_bbox;  // [pool/8] padded bounding box of each block of 8 segments
};

struct TickScratch {
    int* key;       // [slot] cell key, -1 inactive
    int* rank;      // [slot] arrival order inside the cell
    int* cell_count;  // [ncells_padded] -> scanned in place into cell_start
    float2* s_pos; float2* s_vel; float* s_rad; float* s_spd; int* s_slot;
    float2* s_pref; unsigned char* s_alive;
    unsigned char* s_ghost;  // 1: halo / self ghost (multi-GPU): a neighbour candidate only
    int* fb_list;
    int* ev_replan; int* ev_destroyed;
    int ev_cap;     // entries each event list holds; an event beyond it is dropped and counted (C_TOTAL_EV_OVERFLOW)
    unsigned long long* counters;
};

struct GridParams {
    float x0, y0, cell, inv_cell;
    int w, h;
};

// ------------------------------------------------------------------------------------------------
// Cell key of an agent, -1 if it takes no part in the tick.  The reference's own arithmetic can produce a NaN velocity
// (a jammed agent whose LP3D projects two constraints with identical normals, (N_j - N_i).Normalized() = 0: reproduced
// bit for bit by the C oracle, profiles/r02_nan_case.md); from then on the reference is in undefined behaviour (a 0-point
// path is indexed at -1, Simulator.cpp:554).  Here such an agent stays active and keeps its non-finite state, but is
// left out of the neighbour grid - it can be nobody's neighbour anyway, sqDist > EPSILON is false for NaN
// (KDTree.cpp:112) - and is not updated again: otherwise ONE such agent costs every tick the exhaustive scans meant for
// points outside the static grid (measured: tick 0.54 -> 11.7 ms at 1 M agents).
__device__ __forceinline__ int grid_key(const GridParams& gp, float2 p, unsigned* __restrict__ status, unsigned long long* __restrict__ counters, int i) {
    const float fx = (p.x - gp.x0) * gp.inv_cell, fy = (p.y - gp.y0) * gp.inv_cell;
    if (!(fabsf(fx) < CUDART_INF_F) || !(fabsf(fy) < CUDART_INF_F)) {  // NaN or infinite
        status[i] = 256u;  // ECMGPU_ST_NONFINITE
        atomicAdd(&counters[C_TOTAL_NONFINITE], 1ull);
        return -1;
    }
    const int cx = fx >= 0.0f ? (fx < (float)gp.w ? (int)fx : gp.w - 1) : 0;
    const int cy = fy >= 0.0f ? (fy < (float)gp.h ? (int)fy : gp.h - 1) : 0;
    return cy * gp.w + cx;
}

__global__ void __launch_bounds__(256) k_bin_count(int n_slots, const unsigned char* __restrict__ active,
                                                   const float2* __restrict__ pos, GridParams gp, int* __restrict__ cell_count,
                                                   int* __restrict__ key, int* __restrict__ rank, unsigned* __restrict__ status,
                                                   unsigned long long* __restrict__ counters) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_slots) return;
    if (!active[i]) { key[i] = -1; return; }
    const int k = grid_key(gp, pos[i], status, counters, i);
    key[i] = k;
    if (k >= 0) rank[i] = atomicAdd(&cell_count[k], 1);
}

// ------------------------------------------------------------------------------------------------
// Exclusive scan, 4096 elements per block (1024 threads x int4).
constexpr int kScanBlock = 1024;
constexpr int kScanTile = 4096;

__device__ __forceinline__ int block_exclusive_scan(int v, int& total) {
    __shared__ int warp_sums[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
        int s = warp_sums[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += y;
        }
        warp_sums[lane] = s;
    }
    __syncthreads();
    int base = warp > 0 ? warp_sums[warp - 1] : 0;
    total = warp_sums[31];
    __syncthreads();
    return base + x - v;
}

// The scan in ONE launch (decoupled look-back): a reduce / scan-of-sums / add chain of three dependent launches costs
// more than the work itself once the cell table is a rank's share of the world (strips: 20 tiles at 8 x 125 k agents).  Tiles are handed out by an atomic
// ticket, so every predecessor of a tile has started and the look-back cannot wait for a CTA that is not scheduled.
// state[tile] = (epoch << 34) | (flag << 32) | value, flag 1 = the tile's own sum, 2 = the inclusive prefix; entries of an
// earlier launch carry an older epoch and read as "not there yet", so nothing has to be cleared between ticks: the last
// CTA to finish resets the ticket and advances the epoch (ctl[0] epoch, ctl[1] finished tiles, ctl[2] next ticket), and
// the launch replays inside a CUDA graph.
__global__ void __launch_bounds__(kScanBlock) k_scan_onepass(int4* __restrict__ data, int tiles, unsigned long long* __restrict__ state,
                                                             unsigned* __restrict__ ctl) {
    __shared__ int s_tile, s_prefix;
    __shared__ unsigned s_epoch;
    if (threadIdx.x == 0) {
        s_tile = (int)atomicAdd(&ctl[2], 1u);
        s_epoch = *(volatile unsigned*)&ctl[0];
    }
    __syncthreads();
    const int tile = s_tile;
    const unsigned long long ep = (unsigned long long)(s_epoch & 0x3fffffffu) << 34;
    const int idx = tile * kScanBlock + threadIdx.x;
    const int4 v = data[idx];
    int total;
    const int ex = block_exclusive_scan(v.x + v.y + v.z + v.w, total);
    if (threadIdx.x < 32) {  // warp 0 publishes and looks back
        const int lane = threadIdx.x;
        volatile unsigned long long* st = state;
        int prefix = 0;
        if (tile == 0) {
            if (lane == 0) st[0] = ep | (2ull << 32) | (unsigned)total;
        } else {
            if (lane == 0) st[tile] = ep | (1ull << 32) | (unsigned)total;
            int j = tile - 1;  // lane l inspects tile j - l
            for (;;) {
                const int t = j - lane;
                unsigned long long w = 0;
                if (t >= 0) {
                    do { w = st[t]; } while ((w >> 34) != (ep >> 34) || ((w >> 32) & 3ull) == 0ull);
                }
                const unsigned incl = __ballot_sync(0xffffffffu, t >= 0 && ((w >> 32) & 3ull) == 2ull);
                const int stop = incl ? __ffs(incl) - 1 : 31;  // nearest tile with an inclusive prefix, else the whole window
                int val = (t >= 0 && lane <= stop) ? (int)(unsigned)w : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
                prefix += val;
                if (incl || j - 31 <= 0) break;
                j -= 32;
            }
            if (lane == 0) st[tile] = ep | (2ull << 32) | (unsigned)(prefix + total);
        }
        if (lane == 0) s_prefix = prefix;
    }
    __syncthreads();
    const int base = s_prefix + ex;
    int4 o;
    o.x = base; o.y = base + v.x; o.z = o.y + v.y; o.w = o.z + v.z;
    data[idx] = o;
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&ctl[1], 1u) == (unsigned)tiles - 1u) {  // last tile out: ready for the next launch
            ctl[1] = 0u;
            ctl[2] = 0u;
            __threadfence();
            ctl[0] = s_epoch + 1u;
        }
    }
}


// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_scatter(int n_slots, const int* __restrict__ key, const int* __restrict__ rank,
                                                 const int* __restrict__ cell_start, AgentArrays ag, TickScratch sc) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_slots) return;
    int k = key[i];
    if (k < 0) return;
    int p = cell_start[k] + rank[i];
    sc.s_pos[p] = ag.pos[i];
    sc.s_vel[p] = ag.vel[i];
    sc.s_rad[p] = ag.radius[i];
    sc.s_spd[p] = ag.speed[i];
    sc.s_slot[p] = i;
    sc.s_ghost[p] = 0;
}

// ------------------------------------------------------------------------------------------------
struct TickView {
    EcmView ecm;
    ObstView obst;
    BinView bins;
    GridView grid;
    AgentArrays ag;
    TickScratch sc;
    const int* n_sorted_ptr;  // = &cell_start[w*h]: number of agents in the snapshot
    float step;
    int max_ring;
    int record_neighbors;
    // multi-GPU strips (strips.cuh): the grid holds every agent with x in [cover_lo, cover_hi)
    int strips;
    float cover_lo, cover_hi;
    Lp3dQueue lp3d;
};

// UpdateAttractionPointSystem + ApplySteeringForce for snapshot row p; convergent call (whole warp).
__device__ __forceinline__ void attract_agent(const TickView& t, const int p, const int n) {
    const bool valid = p < n && !t.sc.s_ghost[p];
    unsigned st = 0u;
    int slot = 0, np = 2, cell = -2;
    float spd = 0.0f;
    v2 pos = V(0.0f, 0.0f), goal = V(0.0f, 0.0f), attr = V(0.0f, 0.0f);
    const float2* path = t.ag.path_pool;
    const float4* bbox = t.ag.path_bbox;
    bool have = false, alive = true, need_irm = false;
    if (valid) {
        slot = t.sc.s_slot[p];
        pos = t.sc.s_pos[p];
        spd = t.sc.s_spd[p];
        const PathHdr hdr = t.ag.path_hdr[slot];
        path += hdr.off;
        bbox += hdr.off >> 3;
        np = hdr.len;
        goal = V(hdr.gx, hdr.gy);
        // SquareDistance(float,float,float,float) (UtilityFunctions.cpp:43-49)
        const float ddx = pos.x - goal.x, ddy = pos.y - goal.y;
        const float dist = ddx * ddx + ddy * ddy;
        if (dist < 20.0f * 20.0f) {  // arrival radius (Simulator.cpp:543, 557-562)
            attr = goal;
            have = true;
            st |= 4u;
            if (dist < 2.0f * 2.0f) {  // delete distance (Simulator.cpp:542, 564-566)
                alive = false;
                st |= 8u;
                t.ag.active[slot] = 0;
                const unsigned long long e = atomicAdd(&t.sc.counters[C_DESTROYED_N], 1ull);
                if (e < (unsigned long long)t.sc.ev_cap) t.sc.ev_destroyed[e] = slot;
                else atomicAdd(&t.sc.counters[C_TOTAL_EV_OVERFLOW], 1ull);
            }
        } else {
            need_irm = true;
        }
    }
    __syncwarp();
    v2 ap = V(0.0f, 0.0f);  // `Point attractionPoint;` is (0,0) (Simulator.cpp:570)
    const bool ok = find_attraction_point<true>(t.ecm, t.bins, pos, path, bbox, np, goal, ap, cell, need_irm);
    if (valid) {
        if (need_irm) {
            if (ok) {
                attr = ap;
                have = true;
            } else {
                st |= 2u;
                if (cell == -1) st |= 1u;
                if (!t.ag.replan_pending[slot]) {  // one event per request; cleared by ecmgpu_set_path
                    t.ag.replan_pending[slot] = 1;
                    const unsigned long long e = atomicAdd(&t.sc.counters[C_REPLAN_N], 1ull);
                    if (e < (unsigned long long)t.sc.ev_cap) t.sc.ev_replan[e] = slot;
                    else atomicAdd(&t.sc.counters[C_TOTAL_EV_OVERFLOW], 1ull);
                }
            }
        } else {
            cell = -2;
        }
        if (have) t.ag.attraction[slot] = attr;
        else attr = t.ag.attraction[slot];  // previous attraction point is kept (Simulator.cpp:573-587)
        if (alive) {  // ApplySteeringForce (Simulator.cpp:638-657)
            v2 d = vnormalized(vsub(attr, pos));
            v2 pv = vmul(d, spd);
            t.ag.prefvel[slot] = pv;
            t.sc.s_pref[p] = pv;
        }
        t.sc.s_alive[p] = alive ? 1 : 0;
        t.ag.status[slot] = st;
        t.ag.cell[slot] = cell;
    }
    // warp-aggregated statistics
    unsigned m_loc = __ballot_sync(0xffffffffu, (st & 1u) != 0u);
    unsigned m_rep = __ballot_sync(0xffffffffu, (st & 2u) != 0u);
    if ((threadIdx.x & 31) == 0) {
        if (m_loc) atomicAdd(&t.sc.counters[C_TOTAL_LOCFAIL], (unsigned long long)__popc(m_loc));
        if (m_rep) atomicAdd(&t.sc.counters[C_TOTAL_REPLAN], (unsigned long long)__popc(m_rep));
    }
}

// k_attract waits on dependent gathers (header -> block boxes -> polyline block): resident warps hide them
#ifndef ECM_ATTRACT_MINBLOCKS
#define ECM_ATTRACT_MINBLOCKS 9  // 56 registers; 12 / 16 CTAs (40 / 32 registers, spills) measured 1 % / 8 % slower
#endif
__global__ void __launch_bounds__(128, ECM_ATTRACT_MINBLOCKS) k_attract(TickView t) {
    attract_agent(t, blockIdx.x * blockDim.x + threadIdx.x, *t.n_sorted_ptr);
}

// force = v_orca - v (Simulator.cpp:676-677); v += force * (1/mass) * step (Simulator.cpp:622-633);
// p += v * step (Simulator.cpp:603-604)
__device__ __forceinline__ void integrate_agent(const TickView& t, int slot, v2 pos, v2 vel, v2 velocity) {
    const v2 f = V(velocity.x - vel.x, velocity.y - vel.y);
    const float massRecip = 1.0f / 0.8f;
    const v2 nv = V(vel.x + f.x * massRecip * t.step, vel.y + f.y * massRecip * t.step);
    const v2 np = V(pos.x + (nv.x * t.step), pos.y + (nv.y * t.step));
    t.ag.force[slot] = f;
    t.ag.vel[slot] = nv;
    t.ag.pos[slot] = np;
}

// ORCA + integration for one agent whose neighbour list is known.  kSync: convergent call by the
// whole warp, lanes without work pass valid = false.  With kDefer an agent that needs LP3D may be
// parked in t.lp3d instead (k_lp3d integrates it).
template <bool kSync, bool kDefer>
__device__ __forceinline__ unsigned finish_agent(const TickView& t, int p, const Knn& k, bool valid = true) {
    const int slot = valid ? t.sc.s_slot[p] : 0;
    const v2 pos = valid ? t.sc.s_pos[p] : V(0.0f, 0.0f), vel = valid ? t.sc.s_vel[p] : V(0.0f, 0.0f);
    const float rad = valid ? t.sc.s_rad[p] : 0.0f, spd = valid ? t.sc.s_spd[p] : 0.0f;
    const v2 pref = valid ? t.sc.s_pref[p] : V(0.0f, 0.0f);
    const int n_nb = valid ? k.count() : 0;
    unsigned extra = 0u;
    if (valid && t.strips) {  // did the search ball stay inside what this rank can see?
        const float r5 = n_nb == kK ? sqrtf(k.d[kK - 1]) * 1.001f : CUDART_INF_F;
        if (pos.x - r5 < t.cover_lo || pos.x + r5 >= t.cover_hi) extra = 128u;
    }
    OrcaResult r = orca_velocity<kSync, kDefer>(t.obst, t.bins, t.grid, pos, vel, rad, spd, pref, n_nb, k.q, t.step, valid, t.lp3d, p);
    if (!valid) return 0u;
    if (!(r.status & kLp3dDeferred)) integrate_agent(t, slot, pos, vel, r.velocity);
    if (t.record_neighbors) {
#pragma unroll
        for (int j = 0; j < kK; j++) t.ag.nbr[kK * slot + j] = k.q[j] >= 0 ? t.grid.slot_of_row(k.q[j]) : -1;
        t.ag.nbr_cnt[slot] = n_nb;
    }
    return (r.status & ~kLp3dDeferred) | extra;
}

// RandomizedLP3D + integration for the agents k_orca parked, one lane per entry, a warp's 32 entries in step
// (randomized_lp3d_warp; second half of k_fallback).  Convergent: every lane of the warp makes the same trips.
// Measured against every lane running its own program: tick 0.497 -> 0.489 ms on the congested 1 M crowd
// (profiles/r03i_ab_congested_lp3d.jsonl: base vs lanewise), same state bit for bit.
__device__ __forceinline__ void lp3d_parked(const TickView& t) {
    const Lp3dQueue dq = t.lp3d;
    const int n = (int)min(*dq.count, (unsigned long long)dq.cap);
    const int lane = threadIdx.x & 31;
    const int stride = gridDim.x * blockDim.x;
    for (int base = blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < n; base += stride) {
        const int e = base + lane;
        const bool valid = e < n;
        int4 h = make_int4(0, 0, 0, 0);  // (p, nObst, nc, failed)
        float4 o = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (valid) { h = dq.hdr[e]; o = dq.out[e]; }
        Cons cs[kMaxCons], proj[kMaxCons];
        const int trips = warp_max_trip<true>(h.z);
        for (int i = 0; i < trips; i++)
            if (i < h.z) cs[i] = dq.cs[(size_t)i * dq.cap + e];
        v2 out = V(o.x, o.y);
        randomized_lp3d_warp(valid, h.y, cs, h.z, o.z, h.w, out, proj);
        if (valid) integrate_agent(t, t.sc.s_slot[h.x], t.sc.s_pos[h.x], t.sc.s_vel[h.x], out);
    }
}

// 5 CTAs of 256 threads per SM (48 registers): measured 6 % faster than the 64-register build (4 CTAs);
// the few spilled values stay in L1 (A/B on one B200, 1M agents: 0.467 -> 0.439 ms)
#ifndef ECM_ORCA_MINBLOCKS
#define ECM_ORCA_MINBLOCKS 5
#endif
#ifndef ECM_ORCA_BLOCK
#define ECM_ORCA_BLOCK 256
#endif
__device__ __forceinline__ void orca_agent(const TickView& t, const int p, const int n) {
    unsigned st = 0u;
    const bool mine = p < n && !t.sc.s_ghost[p] && t.sc.s_alive[p];
    Knn k;
    bool found = false;
    if (mine) {
        found = knn_grid(k, t.sc.s_pos[p], t.grid, t.max_ring);
        if (!found) {
            st = 32u;
            int e = (int)atomicAdd(&t.sc.counters[C_FALLBACK_N], 1ull);
            t.sc.fb_list[e] = p;
        }
    }
    phase_barrier<true, 0>();  // neighbour search | constraints + LP
    st |= finish_agent<true, true>(t, p, k, mine && found);
    if (st) t.ag.status[t.sc.s_slot[p]] |= st;
    unsigned m_ovf = __ballot_sync(0xffffffffu, (st & 16u) != 0u);
    unsigned m_lp3 = __ballot_sync(0xffffffffu, (st & 64u) != 0u);
    unsigned m_fb = __ballot_sync(0xffffffffu, (st & 32u) != 0u);
    unsigned m_hm = __ballot_sync(0xffffffffu, (st & 128u) != 0u);
    if ((threadIdx.x & 31) == 0) {
        if (m_hm) atomicAdd(&t.sc.counters[C_TOTAL_HALO_MISS], (unsigned long long)__popc(m_hm));
        if (m_ovf) atomicAdd(&t.sc.counters[C_TOTAL_OBST_OVF], (unsigned long long)__popc(m_ovf));
        if (m_lp3) atomicAdd(&t.sc.counters[C_TOTAL_LP3D], (unsigned long long)__popc(m_lp3));
        if (m_fb) atomicAdd(&t.sc.counters[C_TOTAL_FALLBACK], (unsigned long long)__popc(m_fb));
    }
}

__global__ void __launch_bounds__(ECM_ORCA_BLOCK, ECM_ORCA_MINBLOCKS) k_orca(TickView t) {
    orca_agent(t, blockIdx.x * blockDim.x + threadIdx.x, *t.n_sorted_ptr);
}

// Fixed-grid versions for strips with the compact walk (ECMGPU_COMPACT=1): a rank holds n_slots = the GLOBAL crowd but
// its snapshot has only the rows of its share (+ ghosts); launching one CTA per possible row tile would start thousands
// of CTAs that find nothing to do.  One resident wave of CTAs walks the row tiles that exist instead.
__global__ void __launch_bounds__(128, ECM_ATTRACT_MINBLOCKS) k_attract_tiles(TickView t) {
    const int n = *t.n_sorted_ptr;
    for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) attract_agent(t, base + threadIdx.x, n);
}
__global__ void __launch_bounds__(ECM_ORCA_BLOCK, ECM_ORCA_MINBLOCKS) k_orca_tiles(TickView t) {
    const int n = *t.n_sorted_ptr;
    for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {  // uniform per CTA: the phase barriers stay legal
        orca_agent(t, base + threadIdx.x, n);
        __syncthreads();
    }
}

// The stragglers of a tick, one kernel: (a) warp-per-agent exhaustive neighbour search + ORCA for agents
// whose ring budget ran out, (b) LP3D + integration for the agents k_orca parked.
// mode 0: full tick; mode 1: neighbour query only (ecmgpu_find_neighbors)
// 8 CTAs of 128 threads per SM at 64 registers (72 unconstrained: 7 CTAs).  With a short LP3D queue the kernel is as long
// as its slowest warp and the geometry hardly matters (1 M agents from rest: +1 %, congested: -1.5 %); with a long one -
// the 4 M map, where 46 % of the congested crowd needs RandomizedLP3D: 58 k batches for 2 400 warp slots - resident warps
// are throughput: tick 3.09-3.20 -> 2.85 ms (profiles/r04b_ab_c4_fallback_occupancy.jsonl, r04c_ab_*).  12 CTAs at 40
// registers: the same within noise.
#ifndef ECM_FALLBACK_CTAS
#define ECM_FALLBACK_CTAS 8  // CTAs per SM in the tick's launch
#endif
#ifndef ECM_FALLBACK_MINBLOCKS
#define ECM_FALLBACK_MINBLOCKS 8
#endif
__global__ void __launch_bounds__(128, ECM_FALLBACK_MINBLOCKS) k_fallback(TickView t, int mode) {
    const int warps_total = (gridDim.x * blockDim.x) >> 5;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int n = (int)t.sc.counters[C_FALLBACK_N];
    GridView g = t.grid;
    g.n_sorted = *t.n_sorted_ptr;
    for (int i = warp; i < n; i += warps_total) {
        const int p = t.sc.fb_list[i];
        Knn k;
        knn_exhaustive(k, t.sc.s_pos[p], g);
        if (lane == 0) {
            if (mode == 0) {
                unsigned st = finish_agent<false, false>(t, p, k);
                if (st & 16u) atomicAdd(&t.sc.counters[C_TOTAL_OBST_OVF], 1ull);
                if (st & 64u) atomicAdd(&t.sc.counters[C_TOTAL_LP3D], 1ull);
                if (st & 128u) atomicAdd(&t.sc.counters[C_TOTAL_HALO_MISS], 1ull);
                if (st) t.ag.status[t.sc.s_slot[p]] |= st;
            } else {
                const int slot = t.sc.s_slot[p];
                for (int j = 0; j < kK; j++) t.ag.nbr[kK * slot + j] = k.q[j] >= 0 ? t.grid.slot_of_row(k.q[j]) : -1;
                t.ag.nbr_cnt[slot] = k.count();
            }
        }
        __syncwarp();
    }
    if (mode == 0) lp3d_parked(t);
}

// ------------------------------------------------------------------------------------------------
// Query kernels (ecmgpu_locate / ecmgpu_retract / ecmgpu_find_neighbors / ecmgpu_find_obstacles)
__global__ void k_locate(EcmView ecm, BinView bins, int n, const float2* __restrict__ xy, int* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = find_cell<false>(ecm, bins, xy[i]);
}

__global__ void k_retract(EcmView ecm, BinView bins, int n, const float2* __restrict__ xy, unsigned char* __restrict__ ok,
                          float2* __restrict__ out, int* __restrict__ edge) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    v2 p = xy[i], r = V(0.0f, 0.0f);
    int c = find_cell<false>(ecm, bins, p);
    bool good = c >= 0 && retract_in_cell(ecm, c, p, r);
    ok[i] = good ? 1 : 0;
    out[i] = r;
    edge[i] = c >= 0 ? (c >> 1) : -1;
}

__global__ void __launch_bounds__(128) k_knn_query(TickView t) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = *t.n_sorted_ptr;
    if (p >= n || t.sc.s_ghost[p]) return;
    Knn k;
    if (knn_grid(k, t.sc.s_pos[p], t.grid, t.max_ring)) {
        const int slot = t.sc.s_slot[p];
#pragma unroll
        for (int j = 0; j < kK; j++) t.ag.nbr[kK * slot + j] = k.q[j] >= 0 ? t.grid.slot_of_row(k.q[j]) : -1;
        t.ag.nbr_cnt[slot] = k.count();
    } else {
        int e = (int)atomicAdd(&t.sc.counters[C_FALLBACK_N], 1ull);
        t.sc.fb_list[e] = p;
    }
}

// Simulator::ValidSpawnLocation (Simulator.cpp:295-311) for a batch of candidate points on the grid of the CURRENT
// positions: valid iff no active agent centre lies strictly closer than the clearance, with the
// reference's expression fl(fl(dx*dx) + fl(dy*dy)) < fl(c*c), dx = location.x - position.x.  Only the cells the
// clearance box touches (plus one cell of slack for the rounding of the box corners) are scanned; agents beyond the
// grid sit in the border cells of their clamped coordinates, which the clamped box then covers too.
__device__ __forceinline__ bool spawn_location_valid(const GridView& g, v2 loc, float c) {
    const float c2 = c * c;
    int xa, ya, xb, yb;
    g.cell_of(V(loc.x - c, loc.y - c), xa, ya);
    g.cell_of(V(loc.x + c, loc.y + c), xb, yb);
    if (!(c == c) || !(loc.x == loc.x) || !(loc.y == loc.y)) { xa = 0; ya = 0; xb = g.w - 1; yb = g.h - 1; }  // NaN: scan everything
    xa = max(xa - 1, 0); ya = max(ya - 1, 0);
    xb = min(xb + 1, g.w - 1); yb = min(yb + 1, g.h - 1);
    for (int y = ya; y <= yb; y++) {
        const int a = __ldg(&g.cell_start[y * g.w + xa]);
        const int b = __ldg(&g.cell_start[y * g.w + xb + 1]);
        for (int k = a; k < b; k++) {
            const v2 pj = __ldg(&g.s_pos[k]);
            const float dx = loc.x - pj.x, dy = loc.y - pj.y;
            if (dx * dx + dy * dy < c2) return false;
        }
    }
    return true;
}

__global__ void __launch_bounds__(128) k_valid_spawn(GridView g, int n, const float2* __restrict__ xy, const float* __restrict__ clearance,
                                                      unsigned char* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = spawn_location_valid(g, xy[i], clearance[i]) ? 1 : 0;
}

// Simulator::UpdateSpawnAreas' inner loop (Simulator.cpp:501-527) with a counter-based generator instead of C rand():
// request i tries up to max_attempts positions uniform in its spawn box and keeps the first that passes
// ValidSpawnLocation; the goal is uniform in its goal box.  Every draw is a pure function of (seed, counter, i, attempt),
// so a run is reproducible whatever the batching; the STREAM differs from rand()'s, i.e. parity with the reference is
// statistical only.
__device__ __forceinline__ float spawn_u01(unsigned long long seed, unsigned long long counter, unsigned i, unsigned k) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (counter + 1ull) + 0xD1B54A32D192ED03ull * (unsigned long long)i + 0x8CB92BA72F3D8DD7ull * (unsigned long long)k;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;  // splitmix64 finaliser
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (float)(z >> 40) * (1.0f / 16777216.0f);  // 24 bits: [0, 1)
}

__global__ void __launch_bounds__(128) k_draw_spawns(GridView g, int n, const float4* __restrict__ spawn_box, const float4* __restrict__ goal_box,
                                                      const float* __restrict__ clearance, unsigned long long seed, unsigned long long counter,
                                                      int max_attempts, float2* __restrict__ out_start, float2* __restrict__ out_goal,
                                                      unsigned char* __restrict__ out_ok) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 sb = spawn_box[i], gb = goal_box[i];  // xmin ymin xmax ymax
    const float c = clearance[i];
    bool ok = false;
    v2 start = V(sb.x, sb.y);
    for (int a = 0; a < max_attempts && !ok; a++) {
        start = V(sb.x + spawn_u01(seed, counter, (unsigned)i, 2u * a) * (sb.z - sb.x), sb.y + spawn_u01(seed, counter, (unsigned)i, 2u * a + 1u) * (sb.w - sb.y));
        ok = g.n_sorted == 0 || spawn_location_valid(g, start, c);
    }
    out_start[i] = start;
    out_goal[i] = V(gb.x + spawn_u01(seed, counter, (unsigned)i, 0x10000u) * (gb.z - gb.x), gb.y + spawn_u01(seed, counter, (unsigned)i, 0x10001u) * (gb.w - gb.y));
    out_ok[i] = ok ? 1 : 0;
}

__global__ void k_find_obstacles(ObstView ob, BinView bins, float2 pos, float range2, int* out, int cap, int* out_n) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *out_n = find_obstacles<false>(ob, bins, pos, range2, out, cap);
}

// CTA-wide reservation in a list whose length lives in global memory: one atomicAdd per CTA instead of
// one per entry (all entries of a list hit ONE address; 20 k serialised atomics cost ~20 us per tick).
// Returns this thread's index (valid where `want`).  All threads of the CTA must call it.
__device__ __forceinline__ int cta_reserve(bool want, int* counter, int* s_warp /* [33] */) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    const unsigned m = __ballot_sync(0xffffffffu, want);
    if (lane == 0) s_warp[wid] = __popc(m);
    __syncthreads();
    if (wid == 0) {
        int c = lane < nw ? s_warp[lane] : 0, incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        int base = 0;
        if (lane == 0 && total > 0) base = atomicAdd(counter, total);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (lane < nw) s_warp[lane] = base + incl - c;
    }
    __syncthreads();
    const int idx = s_warp[wid] + __popc(m & ((1u << lane) - 1u));
    __syncthreads();  // s_warp is reused by the next reservation
    return idx;
}

// ---- transfers by slot id while the arrays are indexed internally (spatial renumbering, ecmgpu.cu) ----
struct Int5 { int v[5]; };  // one ECMGPU_NEIGHBORS element
template <class T>
__global__ void __launch_bounds__(256) k_gather_slots(int count, int first, const int* __restrict__ int_of, const T* __restrict__ src, T* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) dst[i] = src[int_of[first + i]];
}
template <class T>
__global__ void __launch_bounds__(256) k_scatter_slots(int count, int first, const int* __restrict__ int_of, const T* __restrict__ src, T* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) dst[int_of[first + i]] = src[i];
}

// ecmgpu_update_io with renumbered arrays: the host's slot order in and out in ONE kernel each way (int_of read once)
__global__ void __launch_bounds__(256) k_io_adopt(int count, const int* __restrict__ int_of, const float2* __restrict__ in_pos,
                                                  const float2* __restrict__ in_vel, float2* __restrict__ pos, float2* __restrict__ vel) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const int a = int_of[i];
    if (in_pos) pos[a] = in_pos[i];
    if (in_vel) vel[a] = in_vel[i];
}
__global__ void __launch_bounds__(256) k_io_publish(int count, const int* __restrict__ int_of, const float2* __restrict__ pos,
                                                    const float2* __restrict__ vel, const unsigned char* __restrict__ active,
                                                    float2* __restrict__ out_pos, float2* __restrict__ out_vel, unsigned char* __restrict__ out_active) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const int a = int_of[i];
    if (out_pos) out_pos[i] = pos[a];
    if (out_vel) out_vel[i] = vel[a];
    if (out_active) out_active[i] = active[a];
}

// ---- host I/O as records of owned agents (ecmgpu_update_io_owned) --------------------------------
struct AgentRec {  // == ecmgpu_agent_rec (include/ecm_b200.h)
    int slot;
    float x, y, vx, vy;
};

// Host-provided state: position and velocity of the listed slots, where this handle owns them.
__global__ void __launch_bounds__(256) k_apply_records(int n, const AgentRec* __restrict__ rec, int max_slots, const unsigned char* __restrict__ active,
                                                       float2* __restrict__ pos, float2* __restrict__ vel, const int* __restrict__ int_of) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const AgentRec r = rec[i];
    if (r.slot < 0 || r.slot >= max_slots) return;
    const int a = int_of ? int_of[r.slot] : r.slot;  // records name agents by slot id; arrays are indexed internally
    if (!active[a]) return;
    pos[a] = make_float2(r.x, r.y);
    vel[a] = make_float2(r.vx, r.vy);
}

// Compacts the owned agents into records (one atomic per CTA; ascending slots within a CTA).
constexpr int kCollectBlock = 1024;
__global__ void __launch_bounds__(kCollectBlock) k_collect_owned(int n_slots, const unsigned char* __restrict__ active, const float2* __restrict__ pos,
                                                                 const float2* __restrict__ vel, AgentRec* __restrict__ out, int out_cap, int* __restrict__ count,
                                                                 const int* __restrict__ ext_of) {
    __shared__ int s_warp[33];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool mine = i < n_slots && active[i];
    const int e = cta_reserve(mine, count, s_warp);
    if (!mine || e >= out_cap) return;  // `out` may be the caller's own (pinned) buffer: never past its end; the count tells
    const float2 p = pos[i], v = vel[i];
    AgentRec r;
    r.slot = ext_of ? ext_of[i] : i; r.x = p.x; r.y = p.y; r.vx = v.x; r.vy = v.y;
    out[e] = r;
}

}  // namespace ecm
