// Exact k-nearest neighbours (k = 5) on the per-tick uniform grid, one thread per agent.
//
// Contract ("exact-knn"; replaces KDTree::KNearestAgents, KDTree.cpp:85-202):
//   candidates = agents active at the start of the tick (KDTree::Construct, KDTree.cpp:34-41);
//   sqDist     = fl(fl(dx*dx) + fl(dy*dy)), dx = pos[j].x - pos[i].x   (KDTree.cpp:106-108);
//   keep        sqDist > EPSILON                                        (KDTree.cpp:112, 146);
//   result     = the 5 smallest by (sqDist, slot id) ascending, in that order; count = min(5, kept).
// The reference has no search radius, so rings of cells are added until the 5th distance is
// provably minimal; a thread that exhausts the ring budget reports failure and the exhaustive
// warp-per-agent pass (knn_exhaustive) resolves it.
#pragma once
#include "world.cuh"

namespace ecm {

struct Knn {
    float d[kK];  // ascending by (d, slot)
    int q[kK];    // sorted-snapshot index, -1 = empty
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int j = 0; j < kK; j++) { d[j] = CUDART_INF_F; q[j] = -1; }
    }
    __device__ __forceinline__ int count() const {
        int n = 0;
#pragma unroll
        for (int j = 0; j < kK; j++) n += (q[j] >= 0);
        return n;
    }
    // compare-exchange chain: the candidate sinks to its place, the displaced entry is carried on.
    // Exact distance ties (ordered by slot id) take the out-of-line path: they are rare and keeping
    // them out of the inlined chain keeps the hot loop small (the kernels are I-cache bound).
    __device__ __noinline__ void insert_with_ties(float cd, int cq, const GridView& g) {
        bool placed = false;  // once the candidate sits, the displaced entries just shift down
        for (int j = 0; j < kK; j++) {
            bool less = placed || cd < d[j];
            if (!placed && cd == d[j] && q[j] >= 0) less = g.slot_of_row(cq) < g.slot_of_row(q[j]);  // tie: lower slot id first
            placed = less;
            if (less) {
                float td = d[j]; d[j] = cd; cd = td;
                int tq = q[j]; q[j] = cq; cq = tq;
            }
        }
    }
    __device__ __forceinline__ void insert(float cd, int cq, const GridView& g) {
        const bool tie = (cd == d[0]) | (cd == d[1]) | (cd == d[2]) | (cd == d[3]) | (cd == d[4]);
        if (tie) { insert_with_ties(cd, cq, g); return; }
        bool placed = false;
#pragma unroll
        for (int j = 0; j < kK; j++) {
            const bool less = placed | (cd < d[j]);
            placed = less;
            const float td = less ? d[j] : cd;
            const int tq = less ? q[j] : cq;
            d[j] = less ? cd : d[j];
            q[j] = less ? cq : q[j];
            cd = td;
            cq = tq;
        }
    }
    __device__ __forceinline__ void consider(v2 self, int cand, const GridView& g) {
#ifdef ECM_KNN_STATS  // host test build only (tests/hostdev): counts the candidates visited
        ECM_KNN_STATS;
#endif
        v2 pj = __ldg(&g.s_pos[cand]);
        float dx = pj.x - self.x, dy = pj.y - self.y;
        float dd = dx * dx + dy * dy;
        if (dd > kEpsilon && dd <= d[kK - 1]) insert(dd, cand, g);
    }
};

// Returns true when the result is proven exact within `max_ring` rings.  Ring r is walked as a list
// of row pieces so that there is ONE candidate loop in the code: r = 1: the own row, the row below,
// the row above (3 cells each); r > 1: the two full outer rows, then the two outer cells of every
// row in between.


__device__ __forceinline__ bool knn_grid(Knn& k, v2 self, const GridView& g, int max_ring) {
    int cx, cy;
    g.cell_of(self, cx, cy);
    k.init();
    int r_first = 1;
    for (int r = r_first; r <= max_ring; r++) {
        const int xa = max(cx - r, 0), xb = min(cx + r, g.w - 1);
        const int ya = max(cy - r, 0), yb = min(cy + r, g.h - 1);
        const int pieces = r == 1 ? 3 : 2 + 2 * (2 * r - 1);
        for (int s = 0; s < pieces; s++) {
            int y, x0, x1;
            if (r == 1) {  // own row first: near candidates tighten the 5th distance early
                y = s == 0 ? cy : (s == 1 ? cy - 1 : cy + 1);
                x0 = xa; x1 = xb;
            } else if (s < 2) {
                y = s == 0 ? cy - r : cy + r;
                x0 = xa; x1 = xb;
            } else {
                y = cy - r + 1 + ((s - 2) >> 1);
                x0 = x1 = ((s - 2) & 1) ? cx + r : cx - r;
            }
            if (y < 0 || y >= g.h || x0 < 0 || x1 >= g.w) continue;
            const int a = __ldg(&g.cell_start[y * g.w + x0]);
            const int b = __ldg(&g.cell_start[y * g.w + x1 + 1]);
            for (int c = a; c < b; c++) k.consider(self, c, g);
        }
        // Every agent not scanned yet lies outside the block [xa..xb] x [ya..yb] (agents beyond the
        // grid are clamped into border cells, and a block side on the grid border extends to infinity),
        // hence at distance >= `cover` from `self`.
        float cover = CUDART_INF_F;
        if (xa > 0) cover = fminf(cover, self.x - (g.x0 + (float)xa * g.cell));
        if (xb < g.w - 1) cover = fminf(cover, (g.x0 + (float)(xb + 1) * g.cell) - self.x);
        if (ya > 0) cover = fminf(cover, self.y - (g.y0 + (float)ya * g.cell));
        if (yb < g.h - 1) cover = fminf(cover, (g.y0 + (float)(yb + 1) * g.cell) - self.y);
        if (cover == CUDART_INF_F) return true;  // whole grid scanned
        // 0.999: margin for the rounding of cover and of the squared distances
        if (k.q[kK - 1] >= 0 && cover > 0.0f && k.d[kK - 1] < cover * cover * 0.999f) return true;
    }
    return false;
}


// Exhaustive variant, one WARP per agent: every lane scans a stride of the snapshot, then the 32
// partial lists are merged through shuffles.  All lanes return the same result.
__device__ __forceinline__ void knn_exhaustive(Knn& k, v2 self, const GridView& g) {
    const int lane = threadIdx.x & 31;
    k.init();
    for (int c = lane; c < g.n_sorted; c += 32) k.consider(self, c, g);
    Knn m;
    m.init();
    for (int src = 0; src < 32; src++) {
#pragma unroll
        for (int j = 0; j < kK; j++) {
            float dd = __shfl_sync(0xffffffffu, k.d[j], src);
            int qq = __shfl_sync(0xffffffffu, k.q[j], src);
            if (qq >= 0 && dd <= m.d[kK - 1]) m.insert(dd, qq, g);
        }
    }
    k = m;
}

}  // namespace ecm
