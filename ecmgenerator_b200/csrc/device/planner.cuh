// Batched global path planning on the device: one thread per query, thousands of queries in
// flight.  The caller of the hot path on the spawn / replan side - ECMPathPlanner::FindPath - restated step by step
// so that the polylines equal the reference's (and csrc/host/planner.cpp's) bit for bit:
//   ECMPathPlanner::FindPath                     /root/reference/ECMGenerator/ECMPathPlanner.cpp:22-136
//   AStar::FindPath / ConstructPath              /root/reference/ECMGenerator/AStar.cpp:44-160, 184-212
//   CreateCorridor / ShrinkCorridor              ECMPathPlanner.cpp:146-214
//   TriangulateCorridor / SampleCorridorArc      ECMPathPlanner.cpp:216-278
//   FitPortalRange / Funnel                      ECMPathPlanner.cpp:280-411
//
// A* keeps the reference's open list semantics: a binary heap of vertex ids ordered by the vertex's CURRENT f cost
// (AStarCompare, AStar.h:30-37).  Costs change after a vertex was pushed and while older entries of it are still
// queued, so the heap is not always a valid heap and which entry surfaces next depends on the heap algorithm itself:
// heap_push / heap_pop below restate libstdc++'s std::push_heap / std::pop_heap (__push_heap, __adjust_heap), the
// algorithms behind the reference's std::priority_queue in the oracle's build.
//
// Per-worker scratch lives in global memory (PlanScratch): the A* node records over all vertices (reset through the
// touched list, like the host planner, instead of the reference's O(V) sweep per query, AStar.cpp:162-176), the heap,
// the half-edge path, the portals and the polyline under construction.  Every capacity that an input could exceed is
// checked and reported per query (kPlanOverflow).
//
// What bounds it (profiles/r03b_planner_probe.jsonl, ncu in profiles/r02_experiments.md section 8): ~100 k queries in
// flight, each chasing pointers through its own scratch - far more than any cache holds, so every access of a query is
// a 32-byte DRAM sector of its own (2 MB of DRAM traffic per query).  Hence the layout: ONE 16-byte record per vertex
// (g, f, parent, visited: the four are almost always touched together, as four arrays they were four sectors), and
// capacities sized for the usual query, with the rare query that exceeds them planned again with the full ones
// (ecmgpu_plan_paths, csrc/ecmgpu.cu), so that the scratch budget buys queries in flight instead of head room.
#pragma once
#include "locate.cuh"

namespace ecm {

enum PlanStatus { kPlanOk = 0, kPlanNoPath = 1, kPlanOverflow = 2 };

struct PlanView {
    EcmView ecm;
    BinView bins;
    const float* vert_clear;  // [nV] ECMVertex::clearance
    const int* vert_he;       // [nV] one outgoing half-edge (ECMVertex::half_edge_idx)
    const int* he_next;       // [2 nE] next outgoing half-edge around the same source vertex (ECMHalfEdge::next_idx)
};

struct alignas(16) PlanNode {  // AStarNode (AStar.h:12-28)
    float g, f;   // gCost, fCost: MAX_FLOAT when idle
    int parent;   // parentIndex, nV = INVALID_NODE_INDEX when idle
    int visited;
};

struct PlanScratch {  // worker k owns [k * stride, (k + 1) * stride) of every array
    int n_workers;
    int cap_push, cap_path, cap_portals, cap_out;
    PlanNode* node;          // [workers * nV]
    int* heap;               // [workers * cap_push]
    int* touched;            // [workers * cap_push]
    int* vpath;              // [workers * cap_path] A* vertex path, then reused
    int* epath;              // [workers * cap_path] half-edge path
    float4* portals;         // [workers * cap_portals] (left.x, left.y, right.x, right.y)
    float2* out;             // [workers * cap_out]
};

__device__ __forceinline__ v2 plan_vert(const PlanView& w, int v) { return __ldg(&w.ecm.vert_xy[v]); }
__device__ __forceinline__ int he_target(const PlanView& w, int he) {
    const int2 ev = __ldg(&w.ecm.edge_v[he >> 1]);
    return (he & 1) ? ev.x : ev.y;
}
__device__ __forceinline__ int he_source(const PlanView& w, int he) {
    const int2 ev = __ldg(&w.ecm.edge_v[he >> 1]);
    return (he & 1) ? ev.y : ev.x;
}
// half_edges[0] = {closest_left L0, closest_right R0}, half_edges[1] = {closest_left R1, closest_right L1} (world.cuh)
__device__ __forceinline__ v2 he_closest_left(const PlanView& w, int he) { return __ldg(&w.ecm.edge_cl[4 * (he >> 1) + ((he & 1) ? 3 : 0)]); }
__device__ __forceinline__ v2 he_closest_right(const PlanView& w, int he) { return __ldg(&w.ecm.edge_cl[4 * (he >> 1) + ((he & 1) ? 2 : 1)]); }
// MathUtility::Distance (UtilityFunctions.cpp:17-24)
__device__ __forceinline__ float plan_distance(v2 p1, v2 p2) {
    const float dx = p2.x - p1.x, dy = p2.y - p1.y;
    return sqrtf(dx * dx + dy * dy);
}
// IsLeftOfSegment (UtilityFunctions.cpp:193-196)
__device__ __forceinline__ bool plan_is_left(v2 s0, v2 s1, v2 p) { return (s1.x - s0.x) * (p.y - s0.y) - (s1.y - s0.y) * (p.x - s0.x) > 0.0f; }
// TriangleArea (UtilityFunctions.cpp:203-210)
__device__ __forceinline__ float plan_tri_area(v2 p1, v2 p2, v2 p3) {
    const float ax = p2.x - p1.x, ay = p2.y - p1.y, bx = p3.x - p1.x, by = p3.y - p1.y;
    return bx * ay - ax * by;
}

// std::push_heap after push_back (libstdc++ __push_heap): the new last element climbs while its parent compares
// "greater", i.e. has the larger f cost NOW.
__device__ __forceinline__ void heap_sift_up(int* heap, int hole, int top, int value, const PlanNode* nd) {
    int parent = (hole - 1) / 2;
    const float fv = nd[value].f;  // no cost changes while an element climbs
    while (hole > top && nd[heap[parent]].f > fv) {
        heap[hole] = heap[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    heap[hole] = value;
}
__device__ __forceinline__ void heap_push(int* heap, int& size, int value, const PlanNode* nd) {
    heap_sift_up(heap, size, 0, value, nd);
    size++;
}
// std::pop_heap + pop_back (libstdc++ __pop_heap / __adjust_heap): the hole left by the top sinks to the bottom
// along the children that do NOT compare greater, then the former last element climbs back from there.
__device__ __forceinline__ void heap_pop(int* heap, int& size, const PlanNode* nd) {
    if (size > 1) {
        const int len = size - 1;
        const int value = heap[len];
        heap[len] = heap[0];
        int hole = 0, child = 0;
        while (child < (len - 1) / 2) {
            child = 2 * (child + 1);
            if (nd[heap[child]].f > nd[heap[child - 1]].f) child--;
            heap[hole] = heap[child];
            hole = child;
        }
        if ((len & 1) == 0 && child == (len - 2) / 2) {
            child = 2 * (child + 1);
            heap[hole] = heap[child - 1];
            hole = child - 1;
        }
        heap_sift_up(heap, hole, 0, value, nd);
    }
    size--;
}

// AStar::FindPath (AStar.cpp:44-160) + ConstructPath (:184-212).  vpath receives the vertex path in travel order.
__device__ __forceinline__ int plan_astar(const PlanView& w, int nV, PlanNode* nd, int* heap, int* touched, int cap_push, v2 startLoc, v2 goalLoc,
                                          int startEdge, int goalEdge, float clearance, int* vpath, int cap_path, int& n_path) {
    int n_heap = 0, n_touched = 0;
    const int2 se = __ldg(&w.ecm.edge_v[startEdge]), ge = __ldg(&w.ecm.edge_v[goalEdge]);
    const int sa = se.y, sb = se.x;  // half_edges[0].v_target_idx, half_edges[1].v_target_idx
    const int ga = ge.y, gb = ge.x;
    touched[n_touched++] = sa;
    touched[n_touched++] = sb;
    nd[sa].g = plan_distance(startLoc, plan_vert(w, sa));
    nd[sa].f = nd[sa].g + plan_distance(plan_vert(w, sa), goalLoc);
    nd[sb].g = plan_distance(startLoc, plan_vert(w, sb));
    nd[sb].f = nd[sb].g + plan_distance(plan_vert(w, sb), goalLoc);
    heap_push(heap, n_heap, sa, nd);
    heap_push(heap, n_heap, sb, nd);
    int status = kPlanNoPath;
    n_path = 0;
    while (n_heap > 0) {
        while (n_heap > 0 && nd[heap[0]].visited) heap_pop(heap, n_heap, nd);  // AStar.cpp:82-85
        if (n_heap == 0) break;
        const int cur = heap[0];
        heap_pop(heap, n_heap, nd);
        nd[cur].visited = 1;
        if (__ldg(&w.vert_clear[cur]) < clearance) continue;  // AStar.cpp:99
        if (cur == ga || cur == gb) {
            // reversed: the goal edge's other vertex, the reached one, then the parents back to a start vertex
            int len = 2;
            for (int nxt = nd[cur].parent; nxt < nV; nxt = nd[nxt].parent) len++;
            if (len > cap_path) { status = kPlanOverflow; break; }
            n_path = len;
            vpath[len - 1] = cur == ga ? gb : ga;
            vpath[len - 2] = cur;
            int k = len - 3;
            for (int nxt = nd[cur].parent; nxt < nV; nxt = nd[nxt].parent) vpath[k--] = nxt;
            status = kPlanOk;
            break;
        }
        // neighbour ring (AStar.cpp:104-152); the reference's first walk only comes back to the ring's start
        int he = __ldg(&w.vert_he[cur]);
        const int startNb = he_target(w, he);
        int nb = startNb;
        do {
            he = __ldg(&w.he_next[he]);
            nb = he_target(w, he);
        } while (startNb != nb);
        const v2 cp = plan_vert(w, cur);
        const float gCur = nd[cur].g;
        bool full = false;
        do {
            const PlanNode nn = nd[nb];
            if (!nn.visited) {
                if (n_touched >= cap_push) { full = true; break; }
                heap_push(heap, n_heap, nb, nd);  // pushed with its OLD cost, updated below (AStar.cpp:131-145)
                touched[n_touched++] = nb;
                const v2 np = plan_vert(w, nb);
                const float newG = gCur + plan_distance(cp, np);
                if (newG < nn.g) {
                    const float newF = newG + plan_distance(np, goalLoc);
                    PlanNode upd;
                    upd.g = newG; upd.f = newF; upd.parent = cur; upd.visited = 0;
                    nd[nb] = upd;
                }
            }
            he = __ldg(&w.he_next[he]);
            nb = he_target(w, he);
        } while (nb != startNb);
        if (full) { status = kPlanOverflow; break; }  // the heap / touched list of this pass is full: planned again with room for 2E + 4
    }
    PlanNode idle;
    idle.g = kMaxFloat; idle.f = kMaxFloat; idle.parent = nV; idle.visited = 0;
    for (int k = 0; k < n_touched; k++) nd[touched[k]] = idle;  // CleanRequestData (AStar.cpp:162-176) for what this query touched
    return status;
}

// Corridor entry i (CreateCorridor, ECMPathPlanner.cpp:146-158) and its shrunk bounds (ShrinkCorridor, :160-214).
struct CorridorAt {
    v2 lb, rb, lcb, rcb;
};
__device__ __forceinline__ CorridorAt plan_corridor(const PlanView& w, int he, float clearance) {
    CorridorAt c;
    const int src = he_source(w, he);
    const v2 center = plan_vert(w, src);
    c.lb = he_closest_left(w, he);
    c.rb = he_closest_right(w, he);
    if (__ldg(&w.vert_clear[src]) < clearance) {
        c.lcb = center;
        c.rcb = center;
    } else {
        const v2 ml = vnormalized(vsub(center, c.lb)), mr = vnormalized(vsub(center, c.rb));
        c.lcb = vadd(c.lb, vmul(ml, clearance));
        c.rcb = vadd(c.rb, vmul(mr, clearance));
    }
    return c;
}

__device__ __forceinline__ bool plan_push_portal(float4* portals, int& n, int cap, v2 left, v2 right) {
    if (n >= cap) return false;
    portals[n++] = make_float4(left.x, left.y, right.x, right.y);
    return true;
}

// ECMPathPlanner::SampleCorridorArc (ECMPathPlanner.cpp:256-278)
__device__ __forceinline__ bool plan_arc(float4* portals, int& n, int cap, v2 p1, v2 p2, v2 o1, v2 o2, v2 c, float radius, bool leftArc) {
    const float maxCurveSampleLength = 10.0f;
    bool ok = leftArc ? plan_push_portal(portals, n, cap, p1, o1) : plan_push_portal(portals, n, cap, o1, p1);
    const float edgeLength = vlen(vsub(p2, p1));
    const int numSamples = (int)ceilf(edgeLength / maxCurveSampleLength);
    const float sampleLength = edgeLength / (float)numSamples;
    const v2 edgeDirection = vdiv(vsub(p2, p1), edgeLength);
    for (int i = 0; i < numSamples && ok; i++) {
        v2 p = vadd(p1, vmul(vmul(edgeDirection, sampleLength), (float)i));
        const v2 arcDirection = vnormalized(vsub(p, c));
        p = vadd(c, vmul(arcDirection, radius));
        ok = leftArc ? plan_push_portal(portals, n, cap, p, o2) : plan_push_portal(portals, n, cap, o2, p);
    }
    return ok;
}

// ECMPathPlanner::Funnel (ECMPathPlanner.cpp:323-411) over portals [0, n).  Like the host planner the scan has a step
// budget: the reference never terminates when Point::Approximate(p, p) is false (|coordinate| >= 2048).
__device__ __forceinline__ int plan_funnel(const float4* portals, int n, v2 start, v2 goal, float2* out, int cap_out, int& n_out) {
    long long budget = 64ll * n + 1024;
    v2 apex = start, pl = start, pr = start;
    int leftIdx = 0, rightIdx = 0, apexIdx = 0;
    n_out = 0;
    out[n_out++] = start;
    for (int i = 0; i < n; i++) {
        if (--budget < 0) return kPlanNoPath;
        const float4 po = portals[i];
        const v2 left = V(po.x, po.y), right = V(po.z, po.w);
        if (plan_tri_area(apex, pr, right) <= 0.0f) {
            if (approx(apex, pr) || plan_tri_area(apex, pl, right) > 0.0f) {
                pr = right;
                rightIdx = i;
            } else {
                if (n_out >= cap_out) return kPlanOverflow;
                out[n_out++] = pl;
                apex = pl;
                apexIdx = leftIdx;
                pl = apex; pr = apex;
                leftIdx = apexIdx; rightIdx = apexIdx;
                i = apexIdx;
                continue;
            }
        }
        if (plan_tri_area(apex, pl, left) >= 0.0f) {
            if (approx(apex, pl) || plan_tri_area(apex, pr, left) < 0.0f) {
                pl = left;
                leftIdx = i;
            } else {
                if (n_out >= cap_out) return kPlanOverflow;
                out[n_out++] = pr;
                apex = pr;
                apexIdx = rightIdx;
                pl = apex; pr = apex;
                leftIdx = apexIdx; rightIdx = apexIdx;
                i = apexIdx;
                continue;
            }
        }
    }
    if (!approx(out[n_out - 1], goal)) {
        if (n_out >= cap_out) return kPlanOverflow;
        out[n_out++] = goal;
    }
    return kPlanOk;
}

// ECMPathPlanner::FindPath with preferredAdditionalClearance = 0 (Simulator.cpp:108-112) for worker `k`.
// The polyline is left in the worker's `out` buffer; returns a PlanStatus.
__device__ __forceinline__ int plan_path(const PlanView& w, const PlanScratch& sc, int k, v2 start, v2 goal, float clearance, int& n_out) {
    const int nV = w.ecm.n_vertices;
    float2* out = sc.out + (size_t)k * sc.cap_out;
    n_out = 0;
    // 1. cells and 2. retraction (ECMPathPlanner.cpp:45-70)
    const int cs = find_cell<false>(w.ecm, w.bins, start), cg = find_cell<false>(w.ecm, w.bins, goal);
    if (cs < 0 || cg < 0) return kPlanNoPath;
    v2 rs, rg;
    if (!retract_in_cell(w.ecm, cs, start, rs)) return kPlanNoPath;
    if (!retract_in_cell(w.ecm, cg, goal, rg)) return kPlanNoPath;
    const int startEdge = cs >> 1, goalEdge = cg >> 1;
    if (startEdge == goalEdge) {  // ECMPathPlanner.cpp:74-80
        out[0] = start;
        out[1] = goal;
        n_out = 2;
        return kPlanOk;
    }
    // 3. A* on the medial axis (ECMPathPlanner.cpp:84-89)
    int* vpath = sc.vpath + (size_t)k * sc.cap_path;
    int* epath = sc.epath + (size_t)k * sc.cap_path;
    int n_v = 0;
    int st = plan_astar(w, nV, sc.node + (size_t)k * nV, sc.heap + (size_t)k * sc.cap_push, sc.touched + (size_t)k * sc.cap_push, sc.cap_push, rs, rg,
                        startEdge, goalEdge, clearance, vpath, sc.cap_path, n_v);
    if (st != kPlanOk) return st;
    // half-edge path (ECMPathPlanner.cpp:93-113)
    int m = 0;
    for (int i = 0; i + 1 < n_v; i++) {
        const int i2 = vpath[i + 1];
        int he = __ldg(&w.vert_he[vpath[i]]);
        const int heStart = he;
        do {
            if (he_target(w, he) == i2) { epath[m++] = he; break; }
            he = __ldg(&w.he_next[he]);
        } while (heStart != he);
    }
    if (m == 0) return kPlanNoPath;
    // 4. corridor + 5. portals (ECMPathPlanner.cpp:146-254)
    float4* portals = sc.portals + (size_t)k * sc.cap_portals;
    int np = 0;
    bool ok = true;
    CorridorAt a = plan_corridor(w, epath[0], clearance);
    for (int i = 0; i < m - 1 && ok; i++) {
        const CorridorAt b = plan_corridor(w, epath[i + 1], clearance);
        if (approx(a.lb, b.lb)) {  // LEFT_ARC
            ok = plan_arc(portals, np, sc.cap_portals, a.lcb, b.lcb, a.rcb, b.rcb, a.lb, clearance, true);
        } else if (approx(a.rb, b.rb)) {  // RIGHT_ARC
            ok = plan_arc(portals, np, sc.cap_portals, a.rcb, b.rcb, a.lcb, b.lcb, a.rb, clearance, false);
        } else {  // LINEAR
            ok = plan_push_portal(portals, np, sc.cap_portals, a.lcb, a.rcb) && plan_push_portal(portals, np, sc.cap_portals, b.lcb, a.rcb);
        }
        a = b;
    }
    ok = ok && plan_push_portal(portals, np, sc.cap_portals, a.lcb, a.rcb);
    if (!ok) return kPlanOverflow;
    // FitPortalRange (ECMPathPlanner.cpp:280-320): drop the portals before the first one the start is not left of,
    // and everything from the last one the goal is left of
    int first = 0;
    for (int i = 0; i < np; i++) {
        const float4 po = portals[i];
        if (!plan_is_left(V(po.x, po.y), V(po.z, po.w), start)) { first = i; break; }
    }
    portals += first;
    np -= first;
    int last = np - 1;
    for (int i = np - 1; i >= 0; i--) {
        const float4 po = portals[i];
        if (plan_is_left(V(po.x, po.y), V(po.z, po.w), goal)) { last = i; break; }
    }
    np = max(np - (np - last), 0);
    portals[np++] = make_float4(goal.x, goal.y, goal.x, goal.y);  // overwrites a dropped portal or uses the slot kept free below
    // 7. funnel (ECMPathPlanner.cpp:127-133)
    return plan_funnel(portals, np, start, goal, out, sc.cap_out, n_out);
}

// One thread per worker, queries in a grid-stride loop.  Paths are packed into `pool` through one atomic cursor
// (order of arrival); per query: offset, length (0 = the reference's FindPath returned false) and status.
// subset: the queries to plan (second pass), or nullptr for 0 .. n-1.
constexpr int kPlanBlock = 128;
#ifndef ECM_PLAN_MINBLOCKS
#define ECM_PLAN_MINBLOCKS 6
#endif
constexpr int kPlanThreadsPerSM = kPlanBlock * ECM_PLAN_MINBLOCKS;
__global__ void __launch_bounds__(kPlanBlock, ECM_PLAN_MINBLOCKS) k_plan_paths(PlanView w, PlanScratch sc, int n, const int* __restrict__ subset,
                                                    const float2* __restrict__ start, const float2* __restrict__ goal,
                                                    const float* __restrict__ clearance, int* __restrict__ out_off, int* __restrict__ out_len,
                                                    unsigned char* __restrict__ out_status, float2* __restrict__ pool, int pool_cap, int* cursor) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= sc.n_workers) return;
    const float2* mine = sc.out + (size_t)k * sc.cap_out;
    for (int i = k; i < n; i += sc.n_workers) {
        const int q = subset ? subset[i] : i;
        int len = 0;
        int st = plan_path(w, sc, k, start[q], goal[q], clearance[q], len);
        int off = 0;
        if (st == kPlanOk) {
            off = atomicAdd(cursor, len);
            if (off + len > pool_cap) { st = kPlanOverflow; }
            else for (int j = 0; j < len; j++) pool[off + j] = mine[j];
        }
        if (st != kPlanOk) len = 0;
        out_off[q] = off;
        out_len[q] = len;
        out_status[q] = (unsigned char)st;
    }
}

// Idle state of the A* arrays (AStar::Initialize, AStar.cpp:27-42).
__global__ void __launch_bounds__(256) k_plan_init(PlanScratch sc, int nV) {
    const size_t total = (size_t)sc.n_workers * nV;
    PlanNode idle;
    idle.g = kMaxFloat; idle.f = kMaxFloat; idle.parent = nV; idle.visited = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) sc.node[i] = idle;
}

}  // namespace ecm
