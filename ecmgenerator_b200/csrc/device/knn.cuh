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

// The ordered-insert search: returns true when the result is proven exact within `max_ring` rings.  Ring r is walked as a list
// of row pieces so that there is ONE candidate loop in the code: r = 1: the own row, the row below,
// the row above (3 cells each); r > 1: the two full outer rows, then the two outer cells of every
// row in between.


__device__ __noinline__ bool knn_grid_exact(Knn& k, v2 self, const GridView& g, int max_ring) {
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



// ---- the search the kernels run -------------------------------------------------------------------------------------
// The ordered insert above costs ~30 instructions whenever ANY lane of the warp accepts a candidate, which with 32
// lanes is almost every iteration: ncu attributed 21 % of k_orca's warp instructions to it at 10 of 32 lanes active
// (profiles/r02b_k_orca_by_function.txt).  knn_grid selects branch-free instead.  Every candidate becomes ONE 32-bit key
//     key = (bits of sqDist with the low 13 mantissa bits cleared) | (row offset dy, 4 bits) << 9 | (index in that row, 9 bits)
// which, read as a float, orders like the truncated distance (positive floats order like their bits); the six smallest
// keys are kept by a min / max chain (11 FMNMX per candidate, no branch, every lane busy).  Afterwards the five winners
// are decoded, their EXACT distances recomputed and ordered by (sqDist, slot id).
// Exactness: if the truncated distances of the 5th and the 6th key differ, every candidate left out has a truncated -
// hence an exact - distance strictly above every winner's exact distance: the winners ARE the exact five (ties included).
// If they are equal (relative gap below 2^-10), or a row holds more than 512 candidates, or more than kEncRing rings
// are needed, the ordered-insert search answers instead.  The stop rule uses the upper end of the 5th key's truncation
// interval, so it never stops earlier than the exact rule would.
constexpr int kEncRing = 4;           // rings the key encoding reaches (dy in [-4, 4])
constexpr unsigned kEncLow = 0x1FFFu;  // 13 low bits: 4 (row) + 9 (index)

__device__ __forceinline__ bool knn_grid(Knn& k, v2 self, const GridView& g, int max_ring) {
    int cx, cy;
    g.cell_of(self, cx, cy);
    const float kInf = CUDART_INF_F;
    float t0 = kInf, t1 = kInf, t2 = kInf, t3 = kInf, t4 = kInf, t5 = kInf;
    const int xbase = max(cx - kEncRing, 0);
    bool done = false, fits = true;
    const int rings = min(max_ring, kEncRing);
    for (int r = 1; r <= rings; r++) {
        const int xa = max(cx - r, 0), xb = min(cx + r, g.w - 1);
        const int ya = max(cy - r, 0), yb = min(cy + r, g.h - 1);
        const int pieces = r == 1 ? 3 : 2 + 2 * (2 * r - 1);
        for (int s = 0; s < pieces; s++) {
            int y, x0, x1;
            if (r == 1) {
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
            const int base = __ldg(&g.cell_start[y * g.w + xbase]);
            const int a = __ldg(&g.cell_start[y * g.w + x0]);
            const int b = __ldg(&g.cell_start[y * g.w + x1 + 1]);
            fits = fits && (b - base <= 512);
            const unsigned tag = ((unsigned)(y - cy + kEncRing) << 9) - (unsigned)base;  // + c below = tag | (c - base)
            for (int c = a; c < b; c++) {
#ifdef ECM_KNN_STATS  // host test build only (tests/hostdev): counts the candidates visited
                ECM_KNN_STATS;
#endif
                const v2 pj = __ldg(&g.s_pos[c]);
                const float dx = pj.x - self.x, dy = pj.y - self.y;
                const float dd = dx * dx + dy * dy;
                float x = __uint_as_float((__float_as_uint(dd) & ~kEncLow) + (tag + (unsigned)c));
                x = dd > kEpsilon ? x : kInf;  // also drops NaN
                float lo;
                lo = fminf(t0, x); x = fmaxf(t0, x); t0 = lo;
                lo = fminf(t1, x); x = fmaxf(t1, x); t1 = lo;
                lo = fminf(t2, x); x = fmaxf(t2, x); t2 = lo;
                lo = fminf(t3, x); x = fmaxf(t3, x); t3 = lo;
                lo = fminf(t4, x); x = fmaxf(t4, x); t4 = lo;
                t5 = fminf(t5, x);
            }
        }
        float cover = kInf;
        if (xa > 0) cover = fminf(cover, self.x - (g.x0 + (float)xa * g.cell));
        if (xb < g.w - 1) cover = fminf(cover, (g.x0 + (float)(xb + 1) * g.cell) - self.x);
        if (ya > 0) cover = fminf(cover, self.y - (g.y0 + (float)ya * g.cell));
        if (yb < g.h - 1) cover = fminf(cover, (g.y0 + (float)(yb + 1) * g.cell) - self.y);
        if (cover == kInf) { done = true; break; }  // whole grid scanned
        const float d5_upper = __uint_as_float(__float_as_uint(t4) | kEncLow);  // the 5th exact distance is at most this
        if (t4 < kInf && cover > 0.0f && d5_upper < cover * cover * 0.999f) { done = true; break; }
    }
    const bool ambiguous = t5 < kInf && (__float_as_uint(t4) >> 13) == (__float_as_uint(t5) >> 13);
    if (!done || !fits || ambiguous) {
        // a COPY of the view goes to the out-of-line search: handing out the address of a member of the kernel parameter
        // makes the compiler mirror the whole parameter block (600 bytes) in local memory at kernel entry, for every thread
        // (measured: k_orca's DRAM writes 76 -> 704 MB)
        const GridView gl = g;
        return knn_grid_exact(k, self, gl, max_ring);
    }
    // decode the winners, exact distances, order by (sqDist, slot id)
    k.init();
    const float tk[kK] = {t0, t1, t2, t3, t4};
#pragma unroll
    for (int j = 0; j < kK; j++) {
        if (tk[j] < kInf) {
            const unsigned u = __float_as_uint(tk[j]);
            const int y = cy + (int)((u >> 9) & 15u) - kEncRing;
            const int row = __ldg(&g.cell_start[y * g.w + xbase]) + (int)(u & 511u);
            const v2 pj = __ldg(&g.s_pos[row]);
            const float dx = pj.x - self.x, dy = pj.y - self.y;
            k.d[j] = dx * dx + dy * dy;
            k.q[j] = row;
        }
    }
#pragma unroll
    for (int i = 1; i < kK; i++) {  // insertion sort, 10 compare-exchanges; empty places (+inf, -1) stay last
#pragma unroll
        for (int j = i; j > 0; j--) {
            bool swap = k.d[j] < k.d[j - 1];
            if (k.d[j] == k.d[j - 1] && k.q[j] >= 0 && k.q[j - 1] >= 0) swap = g.slot_of_row(k.q[j]) < g.slot_of_row(k.q[j - 1]);
            const float td = swap ? k.d[j - 1] : k.d[j];
            const int tq = swap ? k.q[j - 1] : k.q[j];
            k.d[j - 1] = swap ? k.d[j] : k.d[j - 1];
            k.q[j - 1] = swap ? k.q[j] : k.q[j - 1];
            k.d[j] = td;
            k.q[j] = tq;
        }
    }
    return true;
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
