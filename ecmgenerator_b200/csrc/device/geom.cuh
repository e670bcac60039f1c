// Device-side 2-D primitives in the reference's operation order.
//
// Bit-exact decisions (cell containment, neighbour ordering) need IEEE binary32 mul/add with no
// FMA contraction and IEEE div/sqrt: this translation unit is compiled with -fmad=false and
// without --use_fast_math (see __graft_entry__.build).  Expression shapes follow
//   /root/reference/ECMGenerator/ECMDataTypes.h:13-104   (Vec2 / Point operators)
//   /root/reference/ECMGenerator/UtilityFunctions.cpp:15-349 (MathUtility)
//   /root/reference/ECMGenerator/Configuration.h:11-15   (EPSILON, MAX_FLOAT)
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

namespace ecm {

constexpr float kEpsilon = 0.0001f;   // Configuration.h:14
constexpr float kMaxFloat = 3.402823466e+38f;  // Utility::MAX_FLOAT (Configuration.h:11)
constexpr float kLookAhead = 10.0f;   // ORCA.h:102-103 (agents and obstacles)
constexpr int kK = 5;                 // Simulator.cpp:55

typedef float2 v2;

// Warp-uniform trip count for loops whose per-lane bound differs: with kSync all 32 lanes of the
// warp iterate max(bound) times and re-converge at the top of every iteration (__syncwarp), so the
// lanes that have work in iteration i do it together instead of drifting apart (independent thread
// scheduling re-converges an early-exit loop only at its exit: ncu showed 3 of 32 lanes active in
// the LP loop, profiles/r01_v1_k_orca_source_hotspots.txt).  kSync requires a convergent call site.
template <bool kSync>
__device__ __forceinline__ int warp_max_trip(int n) {
    if (kSync) return __reduce_max_sync(0xffffffffu, n);
    return n;
}
template <bool kSync>
__device__ __forceinline__ void warp_align() {
    if (kSync) __syncwarp();
}
// Block-wide phase barrier (kSync kernels only): keeps all warps of a CTA in the same phase of a
// long kernel so that the CTA's instruction working set is one phase, not the whole kernel
// (k_orca was instruction-fetch bound: stall_no_instruction 4.9-14.7 per issue, profiles/).
#ifndef ECM_PHASE_BARRIERS
#define ECM_PHASE_BARRIERS 15  // bit i: barrier i of k_orca is there (0 search | 1 obstacles | 2 obstacle half-planes | 3 agent half-planes | LP)
#endif
template <bool kSync, int kId = 0>
__device__ __forceinline__ void phase_barrier() {
    if (kSync && ((ECM_PHASE_BARRIERS >> kId) & 1)) __syncthreads();  // without all four: +4 % (from rest) .. +8 % (congested) per tick, profiles/r02i_ab_*.jsonl
}

__device__ __forceinline__ v2 V(float x, float y) { return make_float2(x, y); }
__device__ __forceinline__ v2 vadd(v2 a, v2 b) { return V(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ v2 vsub(v2 a, v2 b) { return V(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ v2 vmul(v2 a, float s) { return V(a.x * s, a.y * s); }
// IEEE division / sqrt expand to ~10-instruction sequences each; the three helpers built on them are
// kept out of line so that their ~40 call sites do not multiply that code (instruction-cache footprint).
__device__ __noinline__ v2 vdiv(v2 a, float s) { return V(a.x / s, a.y / s); }
__device__ __forceinline__ float vdot(v2 a, v2 b) { return a.x * b.x + a.y * b.y; }   // UtilityFunctions.cpp:172
__device__ __forceinline__ float vdet(v2 a, v2 b) { return a.x * b.y - a.y * b.x; }   // UtilityFunctions.cpp:182
__device__ __forceinline__ float vlen2(v2 a) { return a.x * a.x + a.y * a.y; }        // ECMDataTypes.h:38
__device__ __noinline__ float vlen(v2 a) { return sqrtf(a.x * a.x + a.y * a.y); }  // ECMDataTypes.h:33
__device__ __forceinline__ v2 vright(v2 a) { return V(a.y, -a.x); }                   // UtilityFunctions.cpp:213
__device__ __forceinline__ v2 vleft(v2 a) { return V(-a.y, a.x); }                    // UtilityFunctions.cpp:223
// Vec2::Normalize / Normalized (ECMDataTypes.h:45-60): a zero vector stays zero.
__device__ __noinline__ v2 vnormalized(v2 a) {
    float l = sqrtf(a.x * a.x + a.y * a.y);
    if (l == 0.0f) return a;
    return V(a.x / l, a.y / l);
}
// Arithmetic of the ORCA half-planes and linear programs only (orca.cuh).  Their contract is "new velocities within
// 1e-4 m/s per step" (BASELINE.json north_star), not bit-exactness, so they use the SFU approximations (division 2 ulp,
// square root and reciprocal square root 1-2 ulp) instead of the ~10-instruction IEEE sequences, which were a quarter
// of k_orca's warp instructions (profiles/r01_v8_k_orca_by_function.txt).  Measured on one B200, 1 M agents
// (profiles/r02a_ab_rest.jsonl, r02a_ab_congested.jsonl, r02a_orca_fast_parity.log): tick 0.561 -> 0.497 ms from rest,
// 0.644 -> 0.577 ms congested; worst |dv| per step against the reference 1.2e-7 .. 4.8e-7 m/s (tolerance 1e-4),
// 84-87 % of the velocity rows still bit-identical, 600-tick trajectory RMS 6.5e-7 m.  Everything that feeds a
// bit-exact decision (cells, neighbour order, the obstacle range filter, preferred velocities) stays IEEE.
// -DECM_ORCA_IEEE restores the IEEE sequences: the host-side test builds of this code (tests/hostdev) use it to pin the
// control flow and expression order bit for bit against the reference.
#ifndef ECM_ORCA_IEEE
__device__ __forceinline__ float odiv(float a, float b) { return __fdividef(a, b); }
__device__ __forceinline__ float osqrt(float x) {
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ v2 ovdiv(v2 a, float s) { return V(odiv(a.x, s), odiv(a.y, s)); }
__device__ __forceinline__ float ovlen(v2 a) { return osqrt(__fmaf_rn(a.x, a.x, a.y * a.y)); }
__device__ __forceinline__ v2 ovnormalized(v2 a) {
    const float l2 = __fmaf_rn(a.x, a.x, a.y * a.y);
    if (l2 == 0.0f) return a;
    const float inv = rsqrtf(l2);
    return V(a.x * inv, a.y * inv);
}
// a*b + c*d, a*b - c*d, a*b + c with ONE rounding fewer (this file is compiled with -fmad=false for the bit-exact parts;
// the half-plane arithmetic asks for the contraction explicitly): a third of the instructions of every dot product,
// determinant and point-plus-scaled-direction below.
__device__ __forceinline__ float o2p(float a, float b, float c, float d) { return __fmaf_rn(a, b, c * d); }
__device__ __forceinline__ float o2m(float a, float b, float c, float d) { return __fmaf_rn(a, b, -(c * d)); }
__device__ __forceinline__ float omad(float a, float b, float c) { return __fmaf_rn(a, b, c); }
// sin and cos of atan(r / l), r >= 0, l > 0 (the half-angle of a velocity obstacle, ORCA.cpp:375-379): r and l over the
// hypotenuse - one rsqrt instead of a division, the atanf polynomial and two SFU sine / cosine evaluations.
__device__ __forceinline__ void osincos_atan(float r, float l, float* sn, float* cs) {
    const float inv = rsqrtf(__fmaf_rn(r, r, l * l));
    *sn = r * inv;
    *cs = l * inv;
}
#else
__device__ __forceinline__ float odiv(float a, float b) { return a / b; }
__device__ __forceinline__ float osqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ v2 ovdiv(v2 a, float s) { return vdiv(a, s); }
__device__ __forceinline__ float ovlen(v2 a) { return vlen(a); }
__device__ __forceinline__ v2 ovnormalized(v2 a) { return vnormalized(a); }
__device__ __forceinline__ float o2p(float a, float b, float c, float d) { return a * b + c * d; }
__device__ __forceinline__ float o2m(float a, float b, float c, float d) { return a * b - c * d; }
__device__ __forceinline__ float omad(float a, float b, float c) { return a * b + c; }
__device__ __forceinline__ void osincos_atan(float r, float l, float* sn, float* cs) { sincosf(atanf(r / l), sn, cs); }
#endif
// the vector forms of the above (same expression shapes as vdot / vdet / vlen2 / vadd(p, vmul(d, t)) / vsub(p, vmul(d, t)))
__device__ __forceinline__ float odot(v2 a, v2 b) { return o2p(a.x, b.x, a.y, b.y); }
__device__ __forceinline__ float odet(v2 a, v2 b) { return o2m(a.x, b.y, a.y, b.x); }
__device__ __forceinline__ float olen2(v2 a) { return o2p(a.x, a.x, a.y, a.y); }
__device__ __forceinline__ v2 ovmad(v2 p, v2 d, float t) { return V(omad(d.x, t, p.x), omad(d.y, t, p.y)); }
__device__ __forceinline__ v2 ovmsub(v2 p, v2 d, float t) { return V(omad(-d.x, t, p.x), omad(-d.y, t, p.y)); }
// Point::Approximate (ECMDataTypes.cpp:97-100): open +-EPSILON box.
__device__ __forceinline__ bool approx(v2 a, v2 b) {
    return a.x < (b.x + kEpsilon) && a.x > (b.x - kEpsilon) && a.y < (b.y + kEpsilon) && a.y > (b.y - kEpsilon);
}
// SquareDistance(Point, Point) (UtilityFunctions.cpp:34-41)
__device__ __forceinline__ float sqdist(v2 p1, v2 p2) {
    float dx = p2.x - p1.x, dy = p2.y - p1.y;
    return dx * dx + dy * dy;
}
// GetClosestPointOnSegment (UtilityFunctions.cpp:287-306)
__device__ __forceinline__ v2 closest_on_segment(v2 point, v2 s1, v2 s2) {
    if (approx(s1, s2)) return s1;
    v2 seg = vsub(s2, s1);
    v2 pts = vsub(point, s1);
    float tsq = sqdist(s1, s2);
    float d = (pts.x * seg.x + pts.y * seg.y) / tsq;
    if (d > 1.0f) d = 1.0f;
    if (d < 0.0f) d = 0.0f;
    return V(s1.x + d * seg.x, s1.y + d * seg.y);
}
// RotateVector (UtilityFunctions.cpp:233-242).  sinf/cosf are CUDA's full-range versions (<= 2 ulp),
// glibc's are <= 1 ulp: the only arithmetic on the path that is not bit-reproducible.
__device__ __forceinline__ v2 rotate(v2 v, float rad) {
    float sn, cs;
    sincosf(rad, &sn, &cs);
    return V(v.x * cs - v.y * sn, v.x * sn + v.y * cs);
}

}  // namespace ecm
