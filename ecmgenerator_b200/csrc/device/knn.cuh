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
    // compare-exchange chain: the candidate sinks to its place, the displaced entry is carried on
    __device__ __forceinline__ void insert(float cd, int cq, const int* __restrict__ s_slot) {
#pragma unroll
        for (int j = 0; j < kK; j++) {
            bool less = cd < d[j];
            if (cd == d[j] && q[j] >= 0) less = __ldg(&s_slot[cq]) < __ldg(&s_slot[q[j]]);  // tie: lower slot id first
            if (less) {
                float td = d[j]; d[j] = cd; cd = td;
                int tq = q[j]; q[j] = cq; cq = tq;
            }
        }
    }
    __device__ __forceinline__ void consider(v2 self, int cand, const GridView& g) {
        v2 pj = __ldg(&g.s_pos[cand]);
        float dx = pj.x - self.x, dy = pj.y - self.y;
        float dd = dx * dx + dy * dy;
        if (dd > kEpsilon && dd <= d[kK - 1]) insert(dd, cand, g.s_slot);
    }
};

// Scans the sorted range of cells [cxa, cxb] of row cy.
__device__ __forceinline__ void knn_scan_row(Knn& k, v2 self, const GridView& g, int cy, int cxa, int cxb) {
    int a = __ldg(&g.cell_start[cy * g.w + cxa]);
    int b = __ldg(&g.cell_start[cy * g.w + cxb + 1]);
    for (int c = a; c < b; c++) k.consider(self, c, g);
}

// Returns true when the result is proven exact within `max_ring` rings.
__device__ __forceinline__ bool knn_grid(Knn& k, v2 self, const GridView& g, int max_ring) {
    int cx, cy;
    g.cell_of(self, cx, cy);
    k.init();
    for (int r = 1; r <= max_ring; r++) {
        int xa = max(cx - r, 0), xb = min(cx + r, g.w - 1);
        int ya = max(cy - r, 0), yb = min(cy + r, g.h - 1);
        if (r == 1) {
            for (int y = ya; y <= yb; y++) knn_scan_row(k, self, g, y, xa, xb);
        } else {
            if (cy - r >= 0) knn_scan_row(k, self, g, cy - r, xa, xb);
            if (cy + r < g.h) knn_scan_row(k, self, g, cy + r, xa, xb);
            int y0 = max(cy - r + 1, 0), y1 = min(cy + r - 1, g.h - 1);
            for (int y = y0; y <= y1; y++) {
                if (cx - r >= 0) knn_scan_row(k, self, g, y, cx - r, cx - r);
                if (cx + r < g.w) knn_scan_row(k, self, g, y, cx + r, cx + r);
            }
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
            if (qq >= 0 && dd <= m.d[kK - 1]) m.insert(dd, qq, g.s_slot);
        }
    }
    k = m;
}

}  // namespace ecm
