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
    __device__ __noinline__ void insert_with_ties(float cd, int cq, const int* __restrict__ s_slot) {
        bool placed = false;  // once the candidate sits, the displaced entries just shift down
        for (int j = 0; j < kK; j++) {
            bool less = placed || cd < d[j];
            if (!placed && cd == d[j] && q[j] >= 0) less = __ldg(&s_slot[cq]) < __ldg(&s_slot[q[j]]);  // tie: lower slot id first
            placed = less;
            if (less) {
                float td = d[j]; d[j] = cd; cd = td;
                int tq = q[j]; q[j] = cq; cq = tq;
            }
        }
    }
    __device__ __forceinline__ void insert(float cd, int cq, const int* __restrict__ s_slot) {
        const bool tie = (cd == d[0]) | (cd == d[1]) | (cd == d[2]) | (cd == d[3]) | (cd == d[4]);
        if (tie) { insert_with_ties(cd, cq, s_slot); return; }
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
#ifdef ECM_KNN_BRANCHLESS
        // every lane runs the chain on every candidate (a rejected one carries +inf and changes nothing): with 32
        // lanes almost every iteration has SOME lane inserting, so the branch saved nothing and cost its divergence
        const bool ok = dd > kEpsilon && dd <= d[kK - 1];
        float cd = ok ? dd : CUDART_INF_F;
        const bool tie = ok & ((cd == d[0]) | (cd == d[1]) | (cd == d[2]) | (cd == d[3]) | (cd == d[4]));
        if (tie) { insert_with_ties(cd, cand, g.s_slot); return; }
        int cq = cand;
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
#else
        if (dd > kEpsilon && dd <= d[kK - 1]) insert(dd, cand, g.s_slot);
#endif
    }
};

// Returns true when the result is proven exact within `max_ring` rings.  Ring r is walked as a list
// of row pieces so that there is ONE candidate loop in the code: r = 1: the own row, the row below,
// the row above (3 cells each); r > 1: the two full outer rows, then the two outer cells of every
// row in between.
#ifdef ECM_KNN_PRUNE
// Squared distance from `self` to the column span [x0 .. x1] / to row y of the grid (0 inside).  Border cells
// also hold the agents clamped into them from outside the grid, so they extend to infinity outwards.
__device__ __forceinline__ float knn_span_dx2(const GridView& g, v2 self, int x0, int x1) {
    const float lo = x0 <= 0 ? -CUDART_INF_F : g.x0 + (float)x0 * g.cell;
    const float hi = x1 >= g.w - 1 ? CUDART_INF_F : g.x0 + (float)(x1 + 1) * g.cell;
    const float d = fmaxf(fmaxf(lo - self.x, self.x - hi), 0.0f);
    return d * d;
}
__device__ __forceinline__ float knn_row_dy2(const GridView& g, v2 self, int y) {
    const float lo = y <= 0 ? -CUDART_INF_F : g.y0 + (float)y * g.cell;
    const float hi = y >= g.h - 1 ? CUDART_INF_F : g.y0 + (float)(y + 1) * g.cell;
    const float d = fmaxf(fmaxf(lo - self.y, self.y - hi), 0.0f);
    return d * d;
}
#endif

#ifdef ECM_KNN_TWOPASS
__device__ __forceinline__ bool knn_grid_twopass(Knn& k, v2 self, const GridView& g, int max_ring);
#endif

__device__ __forceinline__ bool knn_grid(Knn& k, v2 self, const GridView& g, int max_ring) {
#ifdef ECM_KNN_TWOPASS
    return knn_grid_twopass(k, self, g, max_ring);
#endif
    int cx, cy;
    g.cell_of(self, cx, cy);
    k.init();
#ifdef ECM_KNN_PRUNE
    // of the two neighbouring rows take the nearer one first: it tightens the 5th distance before the farther one is judged
    const bool low_first = (self.y - (g.y0 + (float)cy * g.cell)) * 2.0f <= g.cell;
#endif
    int r_first = 1;
#ifdef ECM_KNN_FLAT
    // Ring 1 (the 3 x 3 block, where most searches end) as ONE loop over its three row ranges: the six range bounds
    // are loaded up front (independent loads instead of three load -> loop -> load chains), and a warp runs
    // max-over-lanes(n0 + n1 + n2) iterations instead of max(n0) + max(n1) + max(n2).  Same candidates in the same
    // order (own row, row below, row above), hence the same result.
    {
        const int xa = max(cx - 1, 0), xb = min(cx + 1, g.w - 1);
        const int ya = max(cy - 1, 0), yb = min(cy + 1, g.h - 1);
        const int a0 = __ldg(&g.cell_start[cy * g.w + xa]), b0 = __ldg(&g.cell_start[cy * g.w + xb + 1]);
        int a1 = 0, b1 = 0, a2 = 0, b2 = 0;
        if (cy - 1 >= 0) { a1 = __ldg(&g.cell_start[(cy - 1) * g.w + xa]); b1 = __ldg(&g.cell_start[(cy - 1) * g.w + xb + 1]); }
        if (cy + 1 < g.h) { a2 = __ldg(&g.cell_start[(cy + 1) * g.w + xa]); b2 = __ldg(&g.cell_start[(cy + 1) * g.w + xb + 1]); }
        const int n0 = b0 - a0, n01 = n0 + (b1 - a1), total = n01 + (b2 - a2);
        for (int i = 0; i < total; i++) {
            const int c = i < n0 ? a0 + i : (i < n01 ? a1 + (i - n0) : a2 + (i - n01));
            k.consider(self, c, g);
        }
        float cover = CUDART_INF_F;
        if (xa > 0) cover = fminf(cover, self.x - (g.x0 + (float)xa * g.cell));
        if (xb < g.w - 1) cover = fminf(cover, (g.x0 + (float)(xb + 1) * g.cell) - self.x);
        if (ya > 0) cover = fminf(cover, self.y - (g.y0 + (float)ya * g.cell));
        if (yb < g.h - 1) cover = fminf(cover, (g.y0 + (float)(yb + 1) * g.cell) - self.y);
        if (cover == CUDART_INF_F) return true;
        if (k.q[kK - 1] >= 0 && cover > 0.0f && k.d[kK - 1] < cover * cover * 0.999f) return true;
        if (max_ring < 2) return false;
        r_first = 2;
    }
#endif
    for (int r = r_first; r <= max_ring; r++) {
        const int xa = max(cx - r, 0), xb = min(cx + r, g.w - 1);
        const int ya = max(cy - r, 0), yb = min(cy + r, g.h - 1);
        const int pieces = r == 1 ? 3 : 2 + 2 * (2 * r - 1);
        for (int s = 0; s < pieces; s++) {
            int y, x0, x1;
            if (r == 1) {  // own row first: near candidates tighten the 5th distance early
#ifdef ECM_KNN_PRUNE
                y = s == 0 ? cy : ((s == 1) == low_first ? cy - 1 : cy + 1);
#else
                y = s == 0 ? cy : (s == 1 ? cy - 1 : cy + 1);
#endif
                x0 = xa; x1 = xb;
            } else if (s < 2) {
                y = s == 0 ? cy - r : cy + r;
                x0 = xa; x1 = xb;
            } else {
                y = cy - r + 1 + ((s - 2) >> 1);
                x0 = x1 = ((s - 2) & 1) ? cx + r : cx - r;
            }
            if (y < 0 || y >= g.h || x0 < 0 || x1 >= g.w) continue;
#ifdef ECM_KNN_PRUNE
            // A cell farther away than the current 5th distance cannot contribute (distances only shrink, and a tie
            // needs d == d5): drop such end cells of the piece, or the whole piece.  1.001: rounding of both sides.
            if (k.q[kK - 1] >= 0) {
                const float lim = k.d[kK - 1] * 1.001f;
                const float dy2 = knn_row_dy2(g, self, y);
                if (x0 < x1 && dy2 + knn_span_dx2(g, self, x0, x0) > lim) x0++;
                if (x0 < x1 && dy2 + knn_span_dx2(g, self, x1, x1) > lim) x1--;
                if (dy2 + knn_span_dx2(g, self, x0, x1) > lim) continue;
            }
#endif
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

#ifdef ECM_KNN_TWOPASS
// Two-pass variant of knn_grid (same contract, same result).  While the rings expand only the five smallest
// DISTANCES are tracked - a branch-free min / max chain, every lane busy on every candidate; once the 5th distance
// is proven minimal one more sweep over the scanned block collects the candidates with dd <= d5 (normally exactly
// five) and orders them by (distance, slot id).  More than five means exact ties at d5: the ordered insert decides.
__device__ __forceinline__ bool knn_grid_twopass(Knn& k, v2 self, const GridView& g, int max_ring) {
    int cx, cy;
    g.cell_of(self, cx, cy);
    float t0 = CUDART_INF_F, t1 = t0, t2 = t0, t3 = t0, t4 = t0;
    int r = 1;
    bool done = false;
    for (; r <= max_ring; r++) {
        const int xa = max(cx - r, 0), xb = min(cx + r, g.w - 1);
        const int ya = max(cy - r, 0), yb = min(cy + r, g.h - 1);
        const int pieces = r == 1 ? 3 : 2 + 2 * (2 * r - 1);
        for (int s = 0; s < pieces; s++) {
            int y, x0, x1;
            if (r == 1) { y = s == 0 ? cy : (s == 1 ? cy - 1 : cy + 1); x0 = xa; x1 = xb; }
            else if (s < 2) { y = s == 0 ? cy - r : cy + r; x0 = xa; x1 = xb; }
            else { y = cy - r + 1 + ((s - 2) >> 1); x0 = x1 = ((s - 2) & 1) ? cx + r : cx - r; }
            if (y < 0 || y >= g.h || x0 < 0 || x1 >= g.w) continue;
#ifdef ECM_KNN_PRUNE
            if (t4 < CUDART_INF_F) {  // same pruning as knn_grid: cells beyond the current 5th distance cannot contribute
                const float lim = t4 * 1.001f;
                const float dy2 = knn_row_dy2(g, self, y);
                if (x0 < x1 && dy2 + knn_span_dx2(g, self, x0, x0) > lim) x0++;
                if (x0 < x1 && dy2 + knn_span_dx2(g, self, x1, x1) > lim) x1--;
                if (dy2 + knn_span_dx2(g, self, x0, x1) > lim) continue;
            }
#endif
            const int a = __ldg(&g.cell_start[y * g.w + x0]);
            const int b = __ldg(&g.cell_start[y * g.w + x1 + 1]);
            for (int c = a; c < b; c++) {
#ifdef ECM_KNN_STATS
                ECM_KNN_STATS;
#endif
                const v2 pj = __ldg(&g.s_pos[c]);
                const float dx = pj.x - self.x, dy = pj.y - self.y;
                float dd = dx * dx + dy * dy;
                dd = dd > kEpsilon ? dd : CUDART_INF_F;
                float lo;
                lo = fminf(t0, dd); dd = fmaxf(t0, dd); t0 = lo;
                lo = fminf(t1, dd); dd = fmaxf(t1, dd); t1 = lo;
                lo = fminf(t2, dd); dd = fmaxf(t2, dd); t2 = lo;
                lo = fminf(t3, dd); dd = fmaxf(t3, dd); t3 = lo;
                t4 = fminf(t4, dd);
            }
        }
        float cover = CUDART_INF_F;
        if (xa > 0) cover = fminf(cover, self.x - (g.x0 + (float)xa * g.cell));
        if (xb < g.w - 1) cover = fminf(cover, (g.x0 + (float)(xb + 1) * g.cell) - self.x);
        if (ya > 0) cover = fminf(cover, self.y - (g.y0 + (float)ya * g.cell));
        if (yb < g.h - 1) cover = fminf(cover, (g.y0 + (float)(yb + 1) * g.cell) - self.y);
        if (cover == CUDART_INF_F) { done = true; break; }
        if (t4 < CUDART_INF_F && cover > 0.0f && t4 < cover * cover * 0.999f) { done = true; break; }
    }
    if (!done) return false;
    // collecting sweep over the scanned block, row by row (each row is one contiguous range of the snapshot)
    k.init();
    const int xa = max(cx - r, 0), xb = min(cx + r, g.w - 1);
    const int ya = max(cy - r, 0), yb = min(cy + r, g.h - 1);
    int cnt = 0;
    bool overflow = false;
    for (int y = ya; y <= yb; y++) {
        int x0 = xa, x1 = xb;
#ifdef ECM_KNN_PRUNE
        {   // only cells that reach into the final ball can hold a result
            const float lim = t4 * 1.001f;
            const float dy2 = knn_row_dy2(g, self, y);
            if (dy2 > lim) continue;
            while (x0 < x1 && dy2 + knn_span_dx2(g, self, x0, x0) > lim) x0++;
            while (x0 < x1 && dy2 + knn_span_dx2(g, self, x1, x1) > lim) x1--;
        }
#endif
        const int a = __ldg(&g.cell_start[y * g.w + x0]);
        const int b = __ldg(&g.cell_start[y * g.w + x1 + 1]);
        for (int c = a; c < b; c++) {
            const v2 pj = __ldg(&g.s_pos[c]);
            const float dx = pj.x - self.x, dy = pj.y - self.y;
            const float dd = dx * dx + dy * dy;
            if (dd > kEpsilon && dd <= t4) {
                if (cnt < kK) {  // newest first; ordered below
                    k.d[4] = k.d[3]; k.q[4] = k.q[3]; k.d[3] = k.d[2]; k.q[3] = k.q[2];
                    k.d[2] = k.d[1]; k.q[2] = k.q[1]; k.d[1] = k.d[0]; k.q[1] = k.q[0];
                    k.d[0] = dd; k.q[0] = c;
                } else overflow = true;
                cnt++;
            }
        }
    }
    if (overflow) {  // exact ties at the 5th distance: the slot id decides, in the ordered insert
        k.init();
        for (int y = ya; y <= yb; y++) {
            const int a = __ldg(&g.cell_start[y * g.w + xa]);
            const int b = __ldg(&g.cell_start[y * g.w + xb + 1]);
            for (int c = a; c < b; c++) {
                const v2 pj = __ldg(&g.s_pos[c]);
                const float dx = pj.x - self.x, dy = pj.y - self.y;
                const float dd = dx * dx + dy * dy;
                if (dd > kEpsilon && dd <= k.d[kK - 1]) k.insert(dd, c, g.s_slot);
            }
        }
        return true;
    }
    // insertion sort of the <= 5 collected entries by (distance, slot id); empty places hold (+inf, -1) and stay last
    for (int i = 1; i < kK; i++) {
        for (int j = i; j > 0; j--) {
            const bool swap = k.q[j] >= 0 && (k.q[j - 1] < 0 || k.d[j] < k.d[j - 1] ||
                                              (k.d[j] == k.d[j - 1] && __ldg(&g.s_slot[k.q[j]]) < __ldg(&g.s_slot[k.q[j - 1]])));
            if (swap) {
                const float td = k.d[j]; k.d[j] = k.d[j - 1]; k.d[j - 1] = td;
                const int tq = k.q[j]; k.q[j] = k.q[j - 1]; k.q[j - 1] = tq;
            }
        }
    }
    return true;
}
#endif

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
