// ORCA: obstacle gathering, half-plane construction and the 2-D incremental linear programs,
// one thread per agent.  Statement order and float expression shapes follow the reference:
//   Simulator::FindNearestObstacles   /root/reference/ECMAgentSimulator/Simulator.cpp:259-292
//   ORCA::GenerateConstraints         /root/reference/ECMAgentSimulator/ORCA.cpp:60-424
//   ORCA::RandomizedLP                /root/reference/ECMAgentSimulator/ORCA.cpp:428-587
//   ORCA::RandomizedLP3D              /root/reference/ECMAgentSimulator/ORCA.cpp:592-669
#pragma once
#include "knn.cuh"

namespace ecm {

constexpr int kMaxObstNeighbors = 27;             // device cap on FindNearestObstacles' list (status OBST_OVERFLOW beyond)
constexpr int kMaxCons = kMaxObstNeighbors + kK;  // obstacle constraints + 5 agent constraints

// Constraint (ORCA.h:26-80): only m_N and m_PointOnLine are ever read.  .x,.y = normal, .z,.w = point.
typedef float4 Cons;
__device__ __forceinline__ Cons cmake(v2 point, v2 normal) { return make_float4(normal.x, normal.y, point.x, point.y); }
__device__ __forceinline__ v2 cn(const Cons& c) { return V(c.x, c.y); }
__device__ __forceinline__ v2 cp(const Cons& c) { return V(c.z, c.w); }
// Constraint::Contains, methodB (ORCA.h:52)
__device__ __forceinline__ bool ccontains(const Cons& c, v2 p) { return odet(vright(cn(c)), vsub(cp(c), p)) <= 0.0f; }

// The reference filter of Simulator.cpp:271-287 for one segment o -> next[o].
__device__ __forceinline__ bool obstacle_in_range(const ObstView& ob, int o, v2 a, float range2) {
    v2 p = __ldg(&ob.xy[o]), q = __ldg(&ob.xy[__ldg(&ob.next[o])]);
    float sl = vdet(vsub(p, a), vsub(q, p));  // LineLeftDistance (UtilityFunctions.cpp:49-52)
    float sq = (sl * sl) / sqdist(p, q);      // std::powf(s, 2.0f) restated as s*s (DESIGN.md "powf")
    if (sq < range2 && sl < 0.0f) {
        v2 c = closest_on_segment(a, p, q);
        return sqdist(c, a) < range2;
    }
    return false;
}

// FindNearestObstacles through the static bins; ids in (obstacle, vertex) order.  Returns the
// number found (may exceed cap: the excess is dropped and the caller flags OBST_OVERFLOW).
template <bool kSync>
__device__ __forceinline__ int find_obstacles(const ObstView& ob, const BinView& bins, v2 a, float range2, int* out, int cap, bool valid = true) {
    int n = 0;
    const int b = valid ? bins.bin_of(a) : 0;
    int i0 = 0, cnt = 0;
    if (valid && b >= 0) {
        i0 = __ldg(&bins.obst_start[b]);
        cnt = __ldg(&bins.obst_start[b + 1]) - i0;
    }
    const int trips = warp_max_trip<kSync>(cnt);
    for (int k = 0; k < trips; k++) {
        warp_align<kSync>();
        if (k < cnt) {
            int o = __ldg(&bins.obst_items[i0 + k]);
            if (obstacle_in_range(ob, o, a, range2)) { if (n < cap) out[n] = o; n++; }
        }
    }
    if (valid && b < 0 && !bins.closed) {  // outside a static grid that is not known to hold every obstacle's reach: exhaustive scan
        for (int o = 0; o < ob.n; o++)
            if (obstacle_in_range(ob, o, a, range2)) { if (n < cap) out[n] = o; n++; }
    }
    return n;
}

// One obstacle segment -> at most one constraint (ORCA.cpp:70-333).  Returns true if `c` was produced.
__device__ __forceinline__ bool obstacle_constraint(const ObstView& ob, int oL, v2 position, v2 velocity, float clearance, Cons& c) {
    int oR = __ldg(&ob.next[oL]);
    v2 pL = __ldg(&ob.xy[oL]), pR = __ldg(&ob.xy[oR]);
    bool cvxL = __ldg(&ob.convex[oL]) != 0, cvxR = __ldg(&ob.convex[oR]) != 0;
    v2 rp1 = vsub(pL, position), rp2 = vsub(pR, position);
    v2 segDir = vsub(pR, pL);
    const float sp = odiv(odot(vmul(rp1, -1.0f), segDir), olen2(segDir));
    const float distSqLine = olen2(ovmsub(vmul(rp1, -1.0f), segDir, sp));
    const float distSq1 = olen2(rp1), distSq2 = olen2(rp2);
    segDir = __ldg(&ob.dir[oL]);  // = Normalize(segDir)
    const float radiusSq = clearance * clearance;

    if (sp < 0.0f && distSq1 <= radiusSq) {  // collision with the left vertex (ORCA.cpp:92-105)
        if (cvxL) { c = cmake(V(0.0f, 0.0f), ovnormalized(vmul(rp1, -1.0f))); return true; }
        return false;
    } else if (sp > 1.0f && distSq2 <= radiusSq) {  // collision with the right vertex (ORCA.cpp:108-121)
        v2 rnd = __ldg(&ob.dir[oR]);  // = Normalize(next(R).p - R.p)
        if (cvxR && odet(rp2, rnd) >= 0.0f) { c = cmake(V(0.0f, 0.0f), ovnormalized(vmul(rp2, -1.0f))); return true; }
        return false;
    } else if (sp >= 0.0f && sp < 1.0f && distSqLine <= radiusSq) {  // collision with the segment (ORCA.cpp:124-135)
        c = cmake(V(0.0f, 0.0f), vright(segDir));
        return true;
    }

    v2 leftLeg, rightLeg;
    if (sp < 0.0f && distSqLine <= radiusSq) {  // ORCA.cpp:146-169
        if (!cvxL) return false;
        oR = oL; pR = pL; cvxR = cvxL;
        const float leg1 = osqrt(distSq1 - radiusSq);
        leftLeg = ovdiv(V(o2m(rp1.x, leg1, rp1.y, clearance), o2p(rp1.x, clearance, rp1.y, leg1)), distSq1);
        rightLeg = ovdiv(V(o2p(rp1.x, leg1, rp1.y, clearance), o2p(-rp1.x, clearance, rp1.y, leg1)), distSq1);
    } else if (sp > 1.0f && distSqLine <= radiusSq) {  // ORCA.cpp:171-183
        if (!cvxR) return false;
        oL = oR; pL = pR; cvxL = cvxR;
        const float leg2 = osqrt(distSq2 - radiusSq);
        leftLeg = ovdiv(V(o2m(rp2.x, leg2, rp2.y, clearance), o2p(rp2.x, clearance, rp2.y, leg2)), distSq2);
        rightLeg = ovdiv(V(o2p(rp2.x, leg2, rp2.y, clearance), o2p(-rp2.x, clearance, rp2.y, leg2)), distSq2);
    } else {  // ORCA.cpp:186-212
        if (cvxL) {
            const float leg1 = osqrt(distSq1 - radiusSq);
            leftLeg = ovdiv(V(o2m(rp1.x, leg1, rp1.y, clearance), o2p(rp1.x, clearance, rp1.y, leg1)), distSq1);
        } else {
            leftLeg = vmul(segDir, -1.0f);
        }
        if (cvxR) {
            const float leg2 = osqrt(distSq2 - radiusSq);
            rightLeg = ovdiv(V(o2p(rp2.x, leg2, rp2.y, clearance), o2p(-rp2.x, clearance, rp2.y, leg2)), distSq2);
        } else {
            rightLeg = segDir;
        }
    }

    // foreign legs (ORCA.cpp:218-239)
    bool leftForeign = false, rightForeign = false;
    v2 lnd = __ldg(&ob.dir[__ldg(&ob.prev[oL])]);  // = Normalize(L.p - prev(L).p)
    if (cvxL && odet(leftLeg, vmul(lnd, -1.0f)) >= 0.0f) { leftLeg = vmul(lnd, -1.0f); leftForeign = true; }
    v2 rnd = __ldg(&ob.dir[oR]);  // = Normalize(next(R).p - R.p)
    if (cvxR && odet(rightLeg, rnd) <= 0.0f) { rightLeg = rnd; rightForeign = true; }

    const float recip = 1.0f / kLookAhead;  // ORCA.cpp:241
    const v2 leftCutoff = vmul(vsub(pL, position), recip);
    const v2 rightCutoff = vmul(vsub(pR, position), recip);
    const v2 cutoffVec = vsub(rightCutoff, leftCutoff);
    const bool same = (oL == oR);
    const float t = same ? 0.5f : odiv(odot(vsub(velocity, leftCutoff), cutoffVec), olen2(cutoffVec));
    const float tLeft = odot(vsub(velocity, leftCutoff), leftLeg);
    const float tRight = odot(vsub(velocity, rightCutoff), rightLeg);

    if ((t < 0.0f && tLeft < 0.0f) || (same && tLeft < 0.0f && tRight < 0.0f)) {  // ORCA.cpp:259-268
        v2 unitW = ovnormalized(vsub(velocity, leftCutoff));
        c = cmake(ovmad(leftCutoff, vmul(unitW, recip), clearance), unitW);
        return true;
    } else if (t > 1.0f && tRight < 0.0f) {  // ORCA.cpp:270-280
        v2 unitW = ovnormalized(vsub(velocity, rightCutoff));
        c = cmake(ovmad(rightCutoff, vmul(unitW, recip), clearance), unitW);
        return true;
    }
    // ORCA.cpp:284-286
    const float distSqCutoff = (t < 0.0f || t > 1.0f || same) ? CUDART_INF_F : olen2(vsub(velocity, ovmad(leftCutoff, cutoffVec, t)));
    const float distSqLeft = (tLeft < 0.0f) ? CUDART_INF_F : olen2(vsub(velocity, ovmad(leftCutoff, leftLeg, tLeft)));
    const float distSqRight = (tRight < 0.0f) ? CUDART_INF_F : olen2(vsub(velocity, ovmad(rightCutoff, rightLeg, tRight)));

    if (distSqCutoff <= distSqLeft && distSqCutoff <= distSqRight) {  // ORCA.cpp:289-301
        v2 normal = vleft(vmul(segDir, -1.0f));
        c = cmake(ovmad(leftCutoff, vmul(normal, recip), clearance), normal);
        return true;
    } else if (distSqLeft <= distSqRight) {  // ORCA.cpp:303-317
        if (leftForeign) return false;
        v2 normal = vleft(leftLeg);
        c = cmake(ovmad(leftCutoff, vmul(normal, clearance), recip), normal);
        return true;
    }
    if (rightForeign) return false;  // ORCA.cpp:319-332
    v2 normal = vright(rightLeg);
    c = cmake(ovmad(rightCutoff, vmul(normal, clearance), recip), normal);
    return true;
}

// One agent neighbour -> exactly one constraint (ORCA.cpp:339-423).
__device__ __forceinline__ Cons agent_constraint(v2 position, v2 velocity, float clearance, v2 npos, v2 nvel, float nclear, float stepSize) {
    v2 VOPos = V(odiv(npos.x - position.x, kLookAhead), odiv(npos.y - position.y, kLookAhead));
    float VOPosLength = ovlen(VOPos);
    float combinedRadius = nclear + clearance;
    float VORadius = odiv(combinedRadius, kLookAhead);
    v2 relVel = V(velocity.x - nvel.x, velocity.y - nvel.y);
    v2 relPos = V(npos.x - position.x, npos.y - position.y);
    float relPosLength = ovlen(relPos);
    if (relPosLength < combinedRadius) {  // colliding (ORCA.cpp:356-372): uses the sim step, not the look-ahead
        v2 w = vsub(relVel, ovdiv(relPos, stepSize));
        float wLength = ovlen(w);
        v2 unitW = ovdiv(w, wLength);
        v2 U = vmul(unitW, (odiv(combinedRadius, stepSize) - wLength));
        return cmake(ovmad(velocity, U, 0.5f), unitW);
    }
    // the half-angle is atan(VORadius / VOPosLength): atan, not asin (ORCA.cpp:375-376).  RotateVector(VOPos, +a) and
    // RotateVector(VOPos, -a): one sine / cosine pair serves both (sin is odd, cos even, exactly)
    float sn, cs;
    osincos_atan(VORadius, VOPosLength, &sn, &cs);
    v2 VOLeftLeg = V(o2m(VOPos.x, cs, VOPos.y, sn), o2p(VOPos.x, sn, VOPos.y, cs));
    v2 VORightLeg = V(o2p(VOPos.x, cs, VOPos.y, sn), o2m(VOPos.y, cs, VOPos.x, sn));
    v2 base = vsub(VOLeftLeg, VOPos), chk = vsub(relVel, VOPos);
    float sqDistFromCircleCentre = olen2(chk);  // SquareDistance(VOPos, relVel)
    bool liesBelow = odet(base, chk) > 0.0f;  // IsLeftOfVector (UtilityFunctions.cpp:198-201)
    if (liesBelow) {  // ORCA.cpp:385-397
        float distToEdge = VORadius - osqrt(sqDistFromCircleCentre);
        v2 lineNormal = ovnormalized(vsub(relVel, VOPos));
        return cmake(ovmad(velocity, vmul(lineNormal, distToEdge), 0.5f), lineNormal);
    }
    v2 leftPerp = ovdiv(vleft(VOPos), VOPosLength);
    if (odot(leftPerp, relVel) >= 0.0f) {  // closer to the left leg (ORCA.cpp:403-412)
        v2 ln = ovnormalized(VOLeftLeg);
        float l = odot(relVel, ln);  // GetClosestPointOnLineThroughOrigin (UtilityFunctions.cpp:316-320)
        v2 U = vsub(vmul(ln, l), relVel);
        return cmake(ovmad(velocity, U, 0.5f), vleft(ln));
    }
    v2 rn = ovnormalized(VORightLeg);
    float l = odot(relVel, rn);
    v2 U = vsub(vmul(rn, l), relVel);
    return cmake(ovmad(velocity, U, 0.5f), vright(rn));
}

// ORCA::RandomizedLP (ORCA.cpp:428-587).  Returns n on success, else the failing index.  The arithmetic and its
// order are the reference's: constraints are visited in index order, a satisfied one is skipped, a violated one
// projects the optimum onto its line clipped by all earlier lines.
//
// Control flow (kSync: convergent call by the whole warp).  Stepping all lanes through i = 0 .. n-1 together made every
// lane wait through the projection of any lane at every i: with 32 lanes almost every i has SOME lane violated, so the
// warp paid n projections with a third of its lanes working (ncu: SIMT efficiency 0.35 in here).  Instead every lane
// walks at its own pace: (A) skip forward to its next violated constraint - a cheap containment test per step - then
// (B) all lanes that stopped project together, each on its own constraint i (the clip loop runs to the largest i).
// A warp now pays max-over-lanes(violated constraints) projections instead of n.  Per lane the sequence of operations
// is exactly the sequential one, hence bit-identical results.
template <bool kSync>
__device__ __forceinline__ int randomized_lp(const Cons* cs, int n, v2 opt, float maxSpeed, bool useDirOpt, v2& outV) {
    if (useDirOpt) outV = vmul(opt, maxSpeed);
    else if (ovlen(opt) > maxSpeed) outV = vmul(ovnormalized(opt), maxSpeed);
    else outV = opt;
    int result = n;
    int i = 0;
    bool live = n > 0;
    while (kSync ? __any_sync(0xffffffffu, live) : live) {
        // (A) advance to the next violated constraint (ORCA.cpp:477-481: `if (Contains) continue`)
        bool hit = false;
        Cons h = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        while (kSync ? __any_sync(0xffffffffu, live && !hit) : (live && !hit)) {
            if (live && !hit) {
                h = cs[i];
                if (ccontains(h, outV)) { i++; live = i < n; }
                else hit = true;
            }
        }
        // (B) project onto constraint i, clipped by constraints 0 .. i-1
        const v2 dir = vright(cn(h));
        const float dpd = odot(dir, cp(h));
        const float disc = o2p(dpd, dpd, maxSpeed, maxSpeed) - odot(cp(h), cp(h));
        bool bad = disc <= 0.0f;  // `return i` (ORCA.cpp:499-503)
        // `if (disc <= 0) return i; else if (disc > 0) {...}` (ORCA.cpp:499-507) does NEITHER for a NaN discriminant: the
        // constraint is passed over.  It happens: an agent exactly level with an obstacle vertex it touches has sp == 1.0,
        // no collision case applies and the leg is sqrt(distSq - r*r) of a negative number (ORCA.cpp:171-183): a NaN
        // obstacle constraint the reference's LP then ignores (tests/golden/nan_case.npz, found at tick 832 of the 1 M run)
        const bool skip = !bad && !(disc > 0.0f);
        float left = 0.0f, right = 0.0f;
        const bool clip = hit && !bad && !skip;
        if (clip) {
            const float dsq = osqrt(disc);
            left = -dpd - dsq;
            right = -dpd + dsq;
        }
        const int trips = warp_max_trip<kSync>(clip ? i : 0);
        for (int j = 0; j < trips; j++) {
            warp_align<kSync>();
            if (clip && j < i) {
                const Cons hj = cs[j];
                const float den = odet(dir, vright(cn(hj)));
                const float num = odet(vright(cn(hj)), vsub(cp(h), cp(hj)));
                if (fabsf(den) <= kEpsilon) {
                    if (num < 0.0f) bad = true;  // `return i` (ORCA.cpp:526-533)
                } else {
                    const float t = odiv(num, den);
                    if (den >= 0.0f) right = (t < right) ? t : right;  // std::min(right, t)
                    else left = (left < t) ? t : left;                 // std::max(left, t)
                    if (left > right) bad = true;                      // `return i` (ORCA.cpp:546-548); monotone, so order-free
                }
            }
        }
        if (hit) {
            if (bad) {
                result = i;
                live = false;
            } else if (skip) {
                i++;
                live = i < n;
            } else {
                if (useDirOpt) {
                    if (odot(opt, dir) > 0.0f) outV = ovmad(cp(h), dir, right);
                    else outV = ovmad(cp(h), dir, left);
                } else {
                    const float t = odot(dir, vsub(opt, cp(h)));
                    if (t < left) outV = ovmad(cp(h), dir, left);
                    else if (t > right) outV = ovmad(cp(h), dir, right);
                    else outV = ovmad(cp(h), dir, t);
                }
                i++;
                live = i < n;
            }
        }
    }
    return result;
}

// ORCA::RandomizedLP3D (ORCA.cpp:592-669).  `proj` is scratch for the projected constraints.
__device__ __noinline__ void randomized_lp3d(int nObst, const Cons* cs, int total, float maxSpeed, int failed, v2& outV, Cons* proj) {
    float maxPen = 0.0f;
    for (int i = failed; i < total; i++) {
        const Cons ci = cs[i];
        v2 dir = vright(cn(ci));
        if (odet(dir, vsub(cp(ci), outV)) <= maxPen) continue;
        int np = 0;
        for (int k = 0; k < nObst; k++) proj[np++] = cs[k];
        for (int j = nObst; j < i; j++) {
            const Cons cj = cs[j];
            float det = odet(dir, vright(cn(cj)));
            v2 pt;
            if (fabsf(det) <= kEpsilon) {
                if (odot(cn(ci), cn(cj)) > 0.0f) continue;
                pt = vmul(vadd(cp(ci), cp(cj)), 0.5f);
            } else {
                float t = odiv(odet(vright(cn(cj)), vsub(cp(ci), cp(cj))), det);
                pt = ovmad(cp(ci), dir, t);
            }
            proj[np++] = cmake(pt, ovnormalized(vsub(cn(cj), cn(ci))));
        }
        const v2 temp = outV;
        if (randomized_lp<false>(proj, np, cn(ci), maxSpeed, true, outV) < np) outV = temp;
        maxPen = odet(dir, vsub(cp(ci), outV));
    }
}

// The same program for a whole warp of parked agents (k_fallback), one per lane, lanes without an entry pass
// valid = false.  Per lane the operations and their order are those of randomized_lp3d above, hence the same bits; what
// changes is who waits for whom.  Run lane by lane, a warp of 32 different programs executed every lane's constraint
// loop, projection loop and inner 2-D program one after the other (ncu: k_fallback 60 us per tick for the 13 % of the
// congested 1 M crowd that needs it).  Here (A) every lane skips at its own pace to ITS next constraint that is violated
// by more than the penetration reached so far - a cheap test - then (B) the lanes that stopped build their projected
// constraints and run the inner program TOGETHER, through the warp-convergent randomized_lp<true>.
__device__ __forceinline__ void randomized_lp3d_warp(bool valid, int nObst, const Cons* cs, int total, float maxSpeed, int failed, v2& outV, Cons* proj) {
    float maxPen = 0.0f;
    int i = failed;
    bool live = valid && i < total;
    while (__any_sync(0xffffffffu, live)) {
        bool hit = false;
        Cons ci = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        v2 dir = V(0.0f, 0.0f);
        while (__any_sync(0xffffffffu, live && !hit)) {  // (A)
            if (live && !hit) {
                ci = cs[i];
                dir = vright(cn(ci));
                if (odet(dir, vsub(cp(ci), outV)) <= maxPen) { i++; live = i < total; }
                else hit = true;
            }
        }
        // (B) the obstacle constraints as they are, the agent constraints before i projected onto constraint i
        int np = 0;
        {
            const int trips = warp_max_trip<true>(hit ? nObst : 0);
            for (int k = 0; k < trips; k++) {
                warp_align<true>();
                if (hit && k < nObst) proj[np++] = cs[k];
            }
        }
        {
            const int trips = warp_max_trip<true>(hit ? i - nObst : 0);
            for (int jj = 0; jj < trips; jj++) {
                warp_align<true>();
                const int j = nObst + jj;
                if (hit && j < i) {
                    const Cons cj = cs[j];
                    const float det = odet(dir, vright(cn(cj)));
                    v2 pt;
                    bool keep = true;
                    if (fabsf(det) <= kEpsilon) {
                        if (odot(cn(ci), cn(cj)) > 0.0f) keep = false;
                        pt = vmul(vadd(cp(ci), cp(cj)), 0.5f);
                    } else {
                        const float t = odiv(odet(vright(cn(cj)), vsub(cp(ci), cp(cj))), det);
                        pt = ovmad(cp(ci), dir, t);
                    }
                    if (keep) proj[np++] = cmake(pt, ovnormalized(vsub(cn(cj), cn(ci))));
                }
            }
        }
        v2 o = outV;
        const int n_in = hit ? np : 0;
        const int r = randomized_lp<true>(proj, n_in, cn(ci), maxSpeed, true, o);
        if (hit) {
            if (!(r < np)) outV = o;  // an infeasible inner program leaves the velocity as it was (ORCA.cpp:658-664)
            maxPen = odet(dir, vsub(cp(ci), outV));
            i++;
            live = i < total;
        }
    }
}

struct OrcaResult {
    v2 velocity;
    unsigned status;  // ECMGPU_ST_OBST_OVERFLOW | ECMGPU_ST_LP3D | kLp3dDeferred
};

// Agents whose 2-D program is infeasible need RandomizedLP3D: a few per warp in a congested crowd (13 % of
// the agents of the 1 M city crowd after 600 ticks), so run inline almost every warp pays the whole LP3D
// with 4 of 32 lanes working - and k_orca carries its code and its 512-byte scratch array.  k_orca instead
// parks such an agent here - constraints, failing index and the velocity reached so far - and k_lp3d
// finishes the parked agents with every lane busy.  Same functions on the same values: results are
// bit-identical to the inline path (kept for the warp-per-agent fallback kernel).
constexpr unsigned kLp3dDeferred = 0x80000000u;  // internal: never stored in the status array
struct Lp3dQueue {
    int cap;                    // one row per slot
    unsigned long long* count;  // entries this tick (may run past cap: readers clamp)
    int4* hdr;                  // (snapshot row p, nObst, nc, failed)
    float4* out;                // (velocity so far, maxSpeed, unused)
    Cons* cs;                   // constraint i of entry e at [i * cap + e]
};

// ORCA::GetVelocity after the neighbour query (ORCA.cpp:23-56).  nb_q[] are snapshot indices.
// kSync: called by all 32 lanes of the warp (lanes without an agent pass valid = false).
template <bool kSync, bool kDefer>
__device__ __forceinline__ OrcaResult orca_velocity(const ObstView& ob, const BinView& bins, const GridView& g, v2 position, v2 velocity,
                                                    float clearance, float maxSpeed, v2 prefVel, int n_nb, const int* nb_q, float stepSize,
                                                    bool valid, const Lp3dQueue& dq, int p) {
    OrcaResult res;
    res.status = 0u;
    res.velocity = V(0.0f, 0.0f);
    Cons cs[kMaxCons];
    int on[kMaxObstNeighbors];
    if (!valid) n_nb = 0;
    float range = kLookAhead * maxSpeed + clearance;  // ORCA.cpp:27
    int n_on = find_obstacles<kSync>(ob, bins, position, range * range, on, kMaxObstNeighbors, valid);
    if (n_on > kMaxObstNeighbors) { n_on = kMaxObstNeighbors; res.status |= 16u; }
    int nc = 0;
    phase_barrier<kSync, 1>();
    {
        const int trips = warp_max_trip<kSync>(n_on);
        for (int i = 0; i < trips; i++) {
            warp_align<kSync>();
            if (i < n_on) {
                Cons c;
                if (obstacle_constraint(ob, on[i], position, velocity, clearance, c)) cs[nc++] = c;
            }
        }
    }
    const int nObst = nc;
    phase_barrier<kSync, 2>();
    {
        const int trips = warp_max_trip<kSync>(n_nb);
        for (int i = 0; i < trips; i++) {
            warp_align<kSync>();
            if (i < n_nb) {
                int q = nb_q[i];
                cs[nc++] = agent_constraint(position, velocity, clearance, __ldg(&g.s_pos[q]), __ldg(&g.s_vel[q]), __ldg(&g.s_rad[q]), stepSize);
            }
        }
    }
    v2 out = V(0.0f, 0.0f);
    phase_barrier<kSync, 3>();
    int failed = randomized_lp<kSync>(cs, nc, prefVel, maxSpeed, false, out);
    if (failed < nc) {
        res.status |= 64u;
        if constexpr (kDefer) {
            // the queue has a row for every slot and an agent parks at most once per tick: e < cap always holds
            const int e = (int)atomicAdd(dq.count, 1ull);
            if (e < dq.cap) {
                res.status |= kLp3dDeferred;
                dq.hdr[e] = make_int4(p, nObst, nc, failed);
                dq.out[e] = make_float4(out.x, out.y, maxSpeed, 0.0f);
                for (int i = 0; i < nc; i++) dq.cs[(size_t)i * dq.cap + e] = cs[i];
            }
        } else {  // per-lane (k_fallback, queries)
            Cons proj[kMaxCons];
            randomized_lp3d(nObst, cs, nc, maxSpeed, failed, out, proj);
        }
    }
    res.velocity = out;
    return res;
}

}  // namespace ecm
