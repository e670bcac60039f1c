// Faithful KD-tree neighbour mode: the reference's OWN neighbour lists on the GPU.
//
// The default neighbour structure (knn.cuh) implements the exact 5-NN contract of `north_star`.  The reference's
// KDTree::KNearestAgents is not an exact kNN: it prunes against the wrong list entry, can duplicate an id when the
// list fills up, and leaves ids of the PREVIOUS query in the shared output vector when the agent's own node is met
// before five candidates (KDTree.cpp:98-202, ORCA.h:100) - 20-25 % of its lists differ from the true 5-NN.  This
// mode reproduces those lists id for id, so that a whole trajectory can be compared with the UNMODIFIED reference:
//
//   k_kd_init / k_kd_split   median-split tree of the agents active at the start of the tick, level by level:
//                            one segmented sort per level (radix sort of (segment start, coordinate) keys, submitted
//                            by ecmgpu.cu), the lower median of every segment becomes the node
//                            (KDTree::Construct / ConstructRecursive, KDTree.cpp:22-83)
//   k_kd_query               per agent: the reference's pre-order search with its pruning and fill rules,
//                            iteratively (KDTree::KNearestAgents_R, KDTree.cpp:98-202).  List places the search never
//                            wrote hold a TOKEN "place j of the list as the previous query left it"
//   k_kd_resolve             tokens -> ids: the previous query is the one of the preceding live agent in slot order
//                            (ApplyObstacleAvoidanceForce walks the slots upwards, Simulator.cpp:659-686), or the
//                            list carried over from the previous tick for the first one
//   k_kd_cache               keeps the last live agent's list for the next tick (ORCA::m_NeighborCache persists)
//   k_orca_kd                k_orca with these lists instead of the grid search; neighbours are read by SLOT from a
//                            copy of the pre-tick positions / velocities (a stale id may name an agent that is
//                            not active any more; the reference reads its slot all the same)
//
// std::sort leaves the order of agents with EQUAL coordinates implementation-defined.  The tree is unique exactly
// when no agent of a segment ties with the segment's median on the split axis.  Where one does, the reference's
// tree depends on its standard library: libstdc++ (the oracle's build) sorts ranges of up to 16 elements by plain
// insertion - a STABLE sort, whose input order is the parent segment's sorted order - and so does the radix sort
// here (stable, elements start in ascending slot order like KDTree.cpp:34-41), hence such ties resolve identically
// (counted in ecmgpu_stats.kd_small_ties, for information).  A median tie in a LARGER segment goes through
// introsort's unstable partitioning and cannot be followed by a parallel sort: k_kd_split counts those in
// ecmgpu_stats.kd_median_ties - with zero of them the tree equals the reference's.
#pragma once
#include "tick.cuh"

namespace ecm {

constexpr int kKdDead = -1;

// float -> unsigned with the same order (negative values reversed below the positive ones)
__device__ __forceinline__ unsigned kd_orderable(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float kd_from_orderable(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

struct KdBuild {
    int n_slots;                 // elements sorted per level (live agents first, then one dead element per spare place)
    const int* n_active_ptr;     // device: number of agents active at the start of the tick (= rows of the snapshot)
    const int* cell_key;         // [slot] grid cell key of the tick, -1 = not active at the start of the tick (k_bin_count)
    const float2* pos;           // per slot, pre-tick
    float4* tree;                // [cap] (x, y, slot bits, -) per node of the implicit heap (KDTree.cpp:12-20); slot -1 = KDTREE_NULL_NODE
    int cap;
    int* meta;                   // [0] m_MaxDepth (KDTree.cpp:47)
    unsigned long long* ties;    // counter: segments of more than kKdStableRange elements whose median ties on the split axis
    unsigned long long* small_ties;  // counter: the same in smaller segments (resolved like libstdc++'s insertion sort)
};
constexpr int kKdStableRange = 16;  // libstdc++ std::sort: _S_threshold, ranges up to this size are insertion-sorted

// Level 0: element i is slot i (ascending slot order, KDTree.cpp:34-41); the agents active at the start of the tick
// form the root segment [0, n_active) once the first sort (by x) has moved the dead elements behind them.
__global__ void __launch_bounds__(256) k_kd_init(KdBuild b, unsigned long long* __restrict__ keys, int* __restrict__ vals, int* __restrict__ seg_r,
                                                 int* __restrict__ seg_node) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n_slots) return;
    const int n = *b.n_active_ptr;
    if (i == 0) {
        seg_r[0] = n;
        seg_node[0] = 0;
        // m_MaxDepth = ceil(log2(size + 1) - 1) (KDTree.cpp:47): the smallest h with 2^(h+1) >= size + 1
        int h = 0;
        while ((2ll << h) < (long long)n + 1) h++;
        b.meta[0] = h;
    }
    if (b.cell_key[i] >= 0) {
        keys[i] = (unsigned long long)kd_orderable(b.pos[i].x);  // segment start 0 in the high half
        vals[i] = i;
    } else {
        keys[i] = (unsigned long long)(unsigned)b.n_slots << 32;  // behind every segment start
        vals[i] = kKdDead;
    }
}

// After the sort of level `depth`: element i of segment [l, r) (l = high half of its key) is the node if it sits on
// the lower median `mid = l + (r - 1 - l) / 2` (KDTree.cpp:75), goes to the left child [l, mid) if before it and
// to the right child [mid + 1, r) otherwise; its next key carries the other coordinate (KDTree.cpp:65-72).  A
// placed element stays where it is (a segment of its own, never split again), so segment [l, r) IS positions l..r-1.
__global__ void __launch_bounds__(256) k_kd_split(KdBuild b, int depth, const unsigned long long* __restrict__ keys_in, const int* __restrict__ vals_in,
                                                  unsigned long long* __restrict__ keys_out, int* __restrict__ vals_out, const int* __restrict__ seg_r_in,
                                                  const int* __restrict__ seg_node_in, int* __restrict__ seg_r_out, int* __restrict__ seg_node_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n_slots) return;
    const unsigned long long key = keys_in[i];
    const int slot = vals_in[i];
    if (slot == kKdDead) {
        keys_out[i] = (unsigned long long)(unsigned)i << 32;
        vals_out[i] = kKdDead;
        return;
    }
    const int l = (int)(key >> 32);
    const int r = seg_r_in[l], node = seg_node_in[l];
    const int mid = l + (r - 1 - l) / 2;
    if (i == mid) {
        const float2 p = b.pos[slot];
        if (node < b.cap) b.tree[node] = make_float4(p.x, p.y, __int_as_float(slot), 0.0f);
        const float c = kd_from_orderable((unsigned)key);
        const bool tie = (i > l && kd_from_orderable((unsigned)keys_in[i - 1]) == c) || (i + 1 < r && kd_from_orderable((unsigned)keys_in[i + 1]) == c);
        if (tie) atomicAdd(r - l > kKdStableRange ? b.ties : b.small_ties, 1ull);
        keys_out[i] = (unsigned long long)(unsigned)i << 32;
        vals_out[i] = kKdDead;
        return;
    }
    const int nl = i < mid ? l : mid + 1;
    if (i == nl) {  // first element of a child segment publishes its extent
        seg_r_out[nl] = i < mid ? mid : r;
        seg_node_out[nl] = 2 * node + (i < mid ? 1 : 2);
    }
    const float2 p = b.pos[slot];
    const float c = ((depth + 1) & 1) ? p.y : p.x;
    keys_out[i] = ((unsigned long long)(unsigned)nl << 32) | (unsigned long long)kd_orderable(c);
    vals_out[i] = slot;
}

// ---- query --------------------------------------------------------------------------------------------------
struct KdList {
    int ids[kK];
    float dist[kK];
    int found;
};
__device__ __forceinline__ int kd_token(int j) { return -2 - j; }       // place j of the previous query's list
__device__ __forceinline__ bool kd_is_token(int v) { return v <= -2; }
__device__ __forceinline__ int kd_token_place(int v) { return -2 - v; }

// dynamic places of the five-entry lists without local memory
__device__ __forceinline__ int kd_get(const int* a, int j) {
    int v = a[0];
#pragma unroll
    for (int i = 1; i < kK; i++) v = j == i ? a[i] : v;
    return v;
}
template <class T>
__device__ __forceinline__ void kd_set(T* a, int j, T v) {
#pragma unroll
    for (int i = 0; i < kK; i++) a[i] = j == i ? v : a[i];
}

// "Updating the search results" for one tree node (KDTree.cpp:110-171).
__device__ __forceinline__ void kd_visit(KdList& L, float sq, int id) {
    if (L.found < kK && sq > kEpsilon) {  // KDTree.cpp:112-138: fill up
        kd_set(L.ids, L.found, id);
        kd_set(L.dist, L.found, sq);
        L.found++;
        if (L.found == kK) {
            // the largest distance goes to place 0; when it already sat there the new entry overwrites it and
            // appears twice (statement order of KDTree.cpp:133-136)
            float largest = sq;
            int li = kK - 1;
#pragma unroll
            for (int i = 0; i < kK - 1; i++)
                if (L.dist[i] > largest) { largest = L.dist[i]; li = i; }
            L.dist[0] = largest;
            L.ids[0] = kd_get(L.ids, li);
            kd_set(L.dist, li, sq);
            kd_set(L.ids, li, id);
        }
    } else {  // KDTree.cpp:140-170: also taken by the agent's own node (sq <= EPSILON) while the list is still filling
        L.found = kK;
        if (sq < L.dist[0] && sq > kEpsilon) {
            L.dist[0] = sq;
            L.ids[0] = id;
            float largest = sq;
            int li = 0;
#pragma unroll
            for (int i = 1; i < kK; i++)
                if (L.dist[i] > largest) { largest = L.dist[i]; li = i; }
            L.dist[0] = largest;
            L.ids[0] = kd_get(L.ids, li);
            kd_set(L.dist, li, sq);
            kd_set(L.ids, li, id);
        }
    }
}

// KDTree::KNearestAgents (KDTree.cpp:85-96) + KNearestAgents_R (:98-202) without recursion: `near_left` remembers
// for every depth on the current path which child was entered first.
__device__ __forceinline__ void kd_query(const float4* __restrict__ tree, int cap, int max_depth, v2 target, KdList& L) {
#pragma unroll
    for (int j = 0; j < kK; j++) { L.ids[j] = kd_token(j); L.dist[j] = kMaxFloat; }
    L.found = 0;
    int cur = 0, depth = 0;
    unsigned near_left = 0u;
    bool down = true;
    for (;;) {
        if (down) {
            bool present = depth <= max_depth && cur < cap;  // KDTree.cpp:101
            float4 nd = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (present) {
                nd = __ldg(&tree[cur]);
                present = __float_as_int(nd.z) != -1;  // KDTree.cpp:102
            }
            if (present) {
                const float dx = nd.x - target.x, dy = nd.y - target.y;  // KDTree.cpp:106-108
                const float sq = dx * dx + dy * dy;
                kd_visit(L, sq, __float_as_int(nd.z));
                const float cv = (depth & 1) ? nd.y : nd.x, tv = (depth & 1) ? target.y : target.x;  // KDTree.cpp:173-174
                const bool left_first = tv < cv;                                                    // KDTree.cpp:177
                near_left = left_first ? (near_left | (1u << depth)) : (near_left & ~(1u << depth));
                cur = 2 * cur + (left_first ? 1 : 2);
                depth++;
                continue;
            }
            down = false;  // an absent node returns at once
        }
        if (cur == 0) break;
        const int parent = (cur - 1) >> 1;
        depth--;
        const bool was_left = (cur & 1) != 0;
        const bool left_first = ((near_left >> depth) & 1u) != 0u;
        cur = parent;
        if (was_left == left_first) {  // back from the first child: is the other half worth a visit? (KDTree.cpp:181-199)
            const float4 nd = __ldg(&tree[parent]);
            const float cv = (depth & 1) ? nd.y : nd.x, tv = (depth & 1) ? target.y : target.x;
            const float d = tv - cv;
            if (d * d < L.dist[kK - 1]) {  // sic: place k-1, not the largest distance
                cur = 2 * parent + (left_first ? 2 : 1);
                depth++;
                down = true;
            }
        }
    }
}

struct KdQuery {
    const float4* tree;
    int cap;
    const int* meta;
    int* raw;          // [5 * slot] ids or tokens
    int* raw_cnt;      // [slot]
    int* cache;        // [5] the list as the last query of the previous tick left it; zeros at first (ORCA.h:87)
};

// live_only: skip the agents the attraction phase destroyed (a tick); 0: every row of the snapshot (queries)
__global__ void __launch_bounds__(128) k_kd_query(TickView t, KdQuery q, int live_only) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = *t.n_sorted_ptr;
    if (p >= n || t.sc.s_ghost[p] || (live_only && !t.sc.s_alive[p])) return;
    KdList L;
    kd_query(q.tree, q.cap, q.meta[0], t.sc.s_pos[p], L);
    const int slot = t.sc.s_slot[p];
#pragma unroll
    for (int j = 0; j < kK; j++) q.raw[kK * slot + j] = L.ids[j];
    q.raw_cnt[slot] = L.found;
}

// Largest live slot below `i`, -1 if none.  Live = still active after the attraction phase: agents destroyed on
// arrival are skipped by ApplyObstacleAvoidanceForce (Simulator.cpp:664-667).
__device__ __forceinline__ int kd_prev_live(const unsigned char* __restrict__ active, int i) {
    for (i--; i >= 0 && !active[i]; i--) {}
    return i;
}

__global__ void __launch_bounds__(256) k_kd_resolve(int n_slots, const unsigned char* __restrict__ active, KdQuery q, int* __restrict__ nbr,
                                                    int* __restrict__ nbr_cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_slots || !active[i]) return;
    for (int j = 0; j < kK; j++) {
        int v = q.raw[kK * i + j];
        int a = i;
        while (kd_is_token(v)) {  // place `pl` of the list the previous live agent's query left behind
            const int pl = kd_token_place(v);
            a = kd_prev_live(active, a);
            v = a >= 0 ? q.raw[kK * a + pl] : q.cache[pl];
        }
        nbr[kK * i + j] = v;
    }
    nbr_cnt[i] = q.raw_cnt[i];
}

// One CTA: the list of the last live agent becomes the carried-over list of the next tick.
__global__ void __launch_bounds__(256) k_kd_cache(int n_slots, const unsigned char* __restrict__ active, const int* __restrict__ nbr, int* __restrict__ cache) {
    __shared__ int s_last;
    if (threadIdx.x == 0) s_last = -1;
    __syncthreads();
    for (int top = n_slots; top > 0; top -= (int)blockDim.x) {
        const int i = top - 1 - (int)threadIdx.x;
        if (i >= 0 && active[i]) atomicMax(&s_last, i);
        __syncthreads();
        if (s_last >= 0) break;  // uniform: read after the barrier
        __syncthreads();
    }
    const int last = s_last;
    if (last >= 0)
        for (int j = (int)threadIdx.x; j < kK; j += (int)blockDim.x) cache[j] = nbr[kK * last + j];
}

// k_orca with the reference's neighbour lists.  `t.grid.s_pos / s_vel / s_rad` point at SLOT-indexed pre-tick
// copies (set up by ecmgpu.cu), so the ids are used as they are.
__global__ void __launch_bounds__(256, ECM_ORCA_MINBLOCKS) k_orca_kd(TickView t) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = *t.n_sorted_ptr;
    const bool mine = p < n && !t.sc.s_ghost[p] && t.sc.s_alive[p];
    Knn k;
    k.init();
    if (mine) {
        const int slot = t.sc.s_slot[p];
        const int cnt = t.ag.nbr_cnt[slot];
#pragma unroll
        for (int j = 0; j < kK; j++) k.q[j] = j < cnt ? t.ag.nbr[kK * slot + j] : -1;
    }
    __syncthreads();
    const unsigned st = finish_agent<true, true>(t, p, k, mine);
    if (st) t.ag.status[t.sc.s_slot[p]] |= st;
    const unsigned m_ovf = __ballot_sync(0xffffffffu, (st & 16u) != 0u);
    const unsigned m_lp3 = __ballot_sync(0xffffffffu, (st & 64u) != 0u);
    if ((threadIdx.x & 31) == 0) {
        if (m_ovf) atomicAdd(&t.sc.counters[C_TOTAL_OBST_OVF], (unsigned long long)__popc(m_ovf));
        if (m_lp3) atomicAdd(&t.sc.counters[C_TOTAL_LP3D], (unsigned long long)__popc(m_lp3));
    }
}

}  // namespace ecm
