// ECM point location, retraction and the IRM attraction point, one thread per query.
#pragma once
#include "world.cuh"

namespace ecm {

// MathUtility::Contains(Point, vector<Segment>) for the closed chain q0 q1 q2 q3
// (UtilityFunctions.cpp:54-86): even-odd ray cast to +x; a point within EPSILON of any corner is
// "not contained"; strict inequalities, so a point level with a corner is missed by both
// neighbouring cells (reference behaviour, DESIGN.md "location failures").
__device__ __forceinline__ bool contains4(v2 p, v2 q0, v2 q1, v2 q2, v2 q3) {
    bool inside = false;
    v2 a = q0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        v2 b = k == 0 ? q1 : (k == 1 ? q2 : (k == 2 ? q3 : q0));
        if (approx(p, a)) return false;
        if (approx(p, b)) return false;
        if (p.y > fminf(a.y, b.y) && p.y < fmaxf(a.y, b.y) && p.x < fmaxf(a.x, b.x)) {
            float xi = (p.y - a.y) * (b.x - a.x) / (b.y - a.y) + a.x;
            if (a.x == b.x || p.x < xi) inside = !inside;
        }
        a = b;
    }
    return inside;
}

// Cell c = 2*edge + side: polygon [v0, boundary.p0, boundary.p1, v1] with boundary (L0,L1) for the
// left cell and (R0,R1) for the right cell (ECMCellCollection.cpp:24-45, :62-80).
__device__ __forceinline__ bool cell_contains(const EcmView& ecm, int c, v2 p) {
    int e = c >> 1, side = c & 1;
    int2 ev = __ldg(&ecm.edge_v[e]);
    const float2* cl = ecm.edge_cl + 4 * e;
    v2 q0 = __ldg(&ecm.vert_xy[ev.x]);
    v2 q1 = __ldg(&cl[side]);
    v2 q2 = __ldg(&cl[2 + side]);
    v2 q3 = __ldg(&ecm.vert_xy[ev.y]);
    return contains4(p, q0, q1, q2, q3);
}

// ECMCellCollection::PointLocationQueryLinear (ECMCellCollection.cpp:57-90): lowest-index
// containing cell, -1 if none.  The bin list is ascending and a superset of the cells that can
// contain a point of the bin, so the first hit equals the linear scan's.  kSync: see warp_max_trip.
template <bool kSync>
__device__ __forceinline__ int find_cell(const EcmView& ecm, const BinView& bins, v2 p, bool valid = true) {
    int res = -1;
    const bool level = valid && bins.level_hit(p.y);  // exactly level with a cell vertex: the bin lists do not apply (world.cuh)
    const int b = valid && !level ? bins.bin_of(p) : 0;
    int i0 = 0, cnt = 0;
    if (valid && b >= 0) {
        i0 = __ldg(&bins.cell_start[b]);
        cnt = __ldg(&bins.cell_start[b + 1]) - i0;
    }
    const int trips = warp_max_trip<kSync>(cnt);
    for (int k = 0; k < trips; k++) {
        warp_align<kSync>();
        if (k < cnt && res < 0) {
            const int c = __ldg(&bins.cell_items[i0 + k]);
            if (cell_contains(ecm, c, p)) res = c;
        }
    }
    if (level) {  // rare (~1e-6 per query): the cells reaching into the point's bin row, in index order
        const int r = bins.row_of(p);
        if (r >= 0) {
            const int a = __ldg(&bins.row_start[r]), e = __ldg(&bins.row_start[r + 1]);
            for (int k = a; k < e && res < 0; k++) {
                const int c = __ldg(&bins.row_items[k]);
                if (cell_contains(ecm, c, p)) res = c;
            }
            return res;
        }
    }
    if (valid && (b < 0 || level) && !bins.closed) {  // outside a static grid that does not hold every cell: the reference's linear scan
        for (int c = 0; c < 2 * ecm.n_edges && res < 0; c++)
            if (cell_contains(ecm, c, p)) res = c;
    }
    return res;
}

// MathUtility::GetRayToLineSegmentIntersection (UtilityFunctions.cpp:323-349).
// `abs(dot) < 0.000001` compares the float |dot| with a DOUBLE literal; 1e-6 is not a float, so
// the test equals |dot| <= (largest float below 1e-6) = |dot| < 1e-6f rounded up; we keep the
// double compare to be literal about it (one DSETP per agent).
__device__ __forceinline__ bool ray_segment(v2 origin, v2 dir, v2 p1, v2 p2, v2& out) {
    v2 v1 = vsub(origin, p1), vv2 = vsub(p2, p1), v3 = V(-dir.y, dir.x);
    float dot = vdot(vv2, v3);
    if ((double)fabsf(dot) < 0.000001) return false;
    float t1 = vdet(vv2, v1) / dot;
    float t2 = vdot(v1, v3) / dot;
    if (t1 >= 0.0f && (t2 >= 0.0f && t2 <= 1.0f)) {
        out = V(origin.x + dir.x * t1, origin.y + dir.y * t1);
        return true;
    }
    return false;
}

// ECM::RetractPoint (ECM.cpp:20-96) given the located cell.
__device__ __forceinline__ bool retract_in_cell(const EcmView& ecm, int cell, v2 loc, v2& out) {
    int e = cell >> 1;
    int2 ev = __ldg(&ecm.edge_v[e]);
    const float2* cl = ecm.edge_cl + 4 * e;
    v2 p1 = __ldg(&ecm.vert_xy[ev.x]), p2 = __ldg(&ecm.vert_xy[ev.y]);
    // IsLeftOfSegment (UtilityFunctions.cpp:193-196): the side is chosen by the edge, not by the cell
    bool left = (p2.x - p1.x) * (loc.y - p1.y) - (p2.y - p1.y) * (loc.x - p1.x) > 0.0f;
    v2 o1 = __ldg(&cl[left ? 0 : 1]);  // he[0].closest_left  | he[0].closest_right
    v2 o2 = __ldg(&cl[left ? 2 : 3]);  // he[1].closest_right | he[1].closest_left
    v2 ray;
    if (approx(o1, o2)) {  // point obstacle
        ray = vadd(vsub(p1, o1), vsub(p2, o1));
    } else {
        v2 v = vsub(o2, o1);
        ray = left ? V(v.y, -v.x) : V(-v.y, v.x);
    }
    ray = vnormalized(ray);
    return ray_segment(loc, ray, p1, p2, out);
}

// One path segment a -> b against the clearance disk (R, c2): IRMPathFollower.cpp:54-112.
// Returns true if the reference would have written outPoint for this segment; line_hit is set when
// the segment's infinite line meets the disk (`success = true`, IRMPathFollower.cpp:70).
__device__ __forceinline__ bool irm_segment(v2 a, v2 b, v2 R, float c2, v2& out, bool& line_hit) {
    v2 p1 = vsub(a, R), p2 = vsub(b, R);
    v2 ed = vsub(p2, p1);
    float el2 = vlen2(ed);
    float det = vdet(p1, p2);
    float disc = c2 * el2 - det * det;
    if (disc < kEpsilon) return false;
    line_hit = true;
    float dys = ed.y < 0.0f ? -1.0f : 1.0f;
    float sq = sqrtf(disc);
    v2 i1 = V((det * ed.y + dys * ed.x * sq) / el2, (-det * ed.x + fabsf(ed.y) * sq) / el2);
    v2 i2 = V((det * ed.y - dys * ed.x * sq) / el2, (-det * ed.x - fabsf(ed.y) * sq) / el2);
    v2 g1 = vadd(p1, R), g2 = vadd(p2, R), gi1 = vadd(i1, R), gi2 = vadd(i2, R);
    v2 edge = vsub(g2, g1);
    float t1 = vdot(vsub(gi1, g1), edge) / el2;
    float t2 = vdot(vsub(gi2, g1), edge) / el2;
    float maxT = -1.0f;
    bool has = false;
    if (t1 >= 0.0f && t1 <= 1.0f) { out = gi1; maxT = t1; has = true; }
    if (t2 >= 0.0f && t2 <= 1.0f) { if (t2 > maxT) out = gi2; has = true; }
    return has;
}

constexpr int kPathBlock = 8;  // segments per bounding-box block (paths start on multiples of 8 points in the pool)

// IRMPathFollower::FindAttractionPoint (IRMPathFollower.cpp:14-115).  `out` is written exactly
// where the reference writes outPoint; `cell` receives the located cell (-1: location failed).
//
// The reference evaluates every path segment in order and lets later segments overwrite earlier
// ones, i.e. the result is the output of the LAST segment that produces one.  We scan from the end
// and stop at the first producing segment; blocks of 8 segments whose (padded) bounding box the
// clearance disk does not touch cannot produce an output and are skipped unread.  Only if no
// segment produces an output does `success` depend on the line tests of all segments, which a
// plain second pass then evaluates (rare: the agent was pushed off its path).
//
// kSync: called by the whole warp (lanes without a query pass valid = false); the segment scan is
// a warp-synchronous state machine: in every step each still-searching lane evaluates its next
// candidate segment, so the expensive evaluation runs in lock-step instead of lane by lane.
template <bool kSync>
__device__ __forceinline__ bool find_attraction_point(const EcmView& ecm, const BinView& bins, v2 position,
                                                      const float2* __restrict__ path, const float4* __restrict__ bbox, int np, v2 goal,
                                                      v2& out, int& cell, bool valid = true) {
    cell = find_cell<kSync>(ecm, bins, position, valid);
    bool go = valid && cell >= 0;
    bool result = false;
    v2 R = V(0.0f, 0.0f);
    float c2 = 0.0f;
    if (go) go = retract_in_cell(ecm, cell, position, R);
    if (go) {
        const float2* cl = ecm.edge_cl + 4 * (cell >> 1);
        v2 obstA = __ldg(&cl[0]), obstB = __ldg(&cl[2]);  // always the LEFT pair (IRMPathFollower.cpp:31-34)
        v2 closest = closest_on_segment(R, obstA, obstB);
        float clearance = vlen(vsub(R, closest));
        c2 = clearance * clearance;
        if (vlen2(vsub(goal, R)) < c2) {
            out = goal;
            result = true;
            go = false;
        }
    }
    const int nseg = np - 1;
    int b = go ? (nseg + kPathBlock - 1) / kPathBlock : 0;  // blocks still to look at
    int i = 0, i0 = 1;                                      // i < i0: fetch the next overlapping block
    bool found = false, line_hit = false, active = go, cand = false;
    v2 pa = V(0.0f, 0.0f), pb = V(0.0f, 0.0f);
    // Two-speed scan.  Most segments a lane looks at are dismissed by the cheap half of irm_segment - the segment's LINE
    // misses the disk - or because the segment's padded box does not touch the disk (then no intersection point can
    // have t in [0, 1]: the padding is the block boxes', far above the rounding of the intersection arithmetic); only
    // the survivors need the square root and six divisions.  Stepping all lanes segment by segment made every lane wait
    // through the expensive half of any lane (ncu: 7.7 of 32 lanes active in it).  So (A) every lane skips ahead at its
    // own pace to its next surviving segment, then (B) the lanes that have one evaluate it together.
    const float kPad = 0.05f;
    while (kSync ? __any_sync(0xffffffffu, active) : active) {
        while (kSync ? __any_sync(0xffffffffu, active && !cand) : (active && !cand)) {
            if (active && !cand) {
                if (i < i0) {
                    for (;;) {
                        if (--b < 0) break;
                        const float4 bb = __ldg(&bbox[b]);
                        const float dx = fmaxf(fmaxf(bb.x - R.x, R.x - bb.z), 0.0f), dy = fmaxf(fmaxf(bb.y - R.y, R.y - bb.w), 0.0f);
                        if (!(dx * dx + dy * dy > c2)) break;  // bbox is padded on the host: conservative
                    }
                    if (b < 0) {
                        active = false;
                    } else {
                        i0 = b * kPathBlock;
                        i = min(i0 + kPathBlock, nseg) - 1;
                        pb = path[i + 1];
                    }
                }
                if (active) {
                    pa = path[i];
                    const v2 p1 = vsub(pa, R), p2 = vsub(pb, R);
                    const float el2 = vlen2(vsub(p2, p1)), det = vdet(p1, p2);
                    bool want = false;
                    if (!(c2 * el2 - det * det < kEpsilon)) {  // the line meets the disk (IRMPathFollower.cpp:66-70)
                        line_hit = true;
                        const float dx = fmaxf(fmaxf(fminf(pa.x, pb.x) - kPad - R.x, R.x - (fmaxf(pa.x, pb.x) + kPad)), 0.0f);
                        const float dy = fmaxf(fmaxf(fminf(pa.y, pb.y) - kPad - R.y, R.y - (fmaxf(pa.y, pb.y) + kPad)), 0.0f);
                        want = !(dx * dx + dy * dy > c2);
                    }
                    if (want) cand = true;
                    else { pb = pa; i--; }
                }
            }
        }
        if (cand) {
            if (irm_segment(pa, pb, R, c2, out, line_hit)) { found = true; active = false; }
            pb = pa;
            i--;
            cand = false;
        }
    }
    if (go) {
        if (found || line_hit) {
            result = true;  // line_hit without a point: outPoint stays untouched (Simulator.cpp:570-577)
        } else {            // no output anywhere: `success` = any segment line meets the disk, over ALL segments
            v2 pa = path[0];
            for (int k = 0; k < nseg && !result; k++) {
                const v2 pn = path[k + 1];
                const v2 p1 = vsub(pa, R), p2 = vsub(pn, R);
                const float el2 = vlen2(vsub(p2, p1)), det = vdet(p1, p2);
                if (!(c2 * el2 - det * det < kEpsilon)) result = true;
                pa = pn;
            }
        }
    }
    return result;
}

}  // namespace ecm
