/* TEST INFRASTRUCTURE - NOT PRODUCT CODE.  See ecm_oracle.h.
 *
 * Plain-C restatement of the reference's per-tick agent update.  Paths below are relative to
 * /root/reference.  Arithmetic is IEEE binary32 in the reference's operation order; compile with
 * -ffp-contract=off.  "MSVC float overloads" (SURVEY.md H5): unqualified abs/sqrt/atan/cos/sin on
 * float arguments are the float functions.
 */
#define _GNU_SOURCE
#include "ecm_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define EO_EPSILON 0.0001f   /* ECMGenerator/Configuration.h:14 */
#define EO_MAX_FLOAT FLT_MAX /* ECMGenerator/Configuration.h:11 */
#define EO_K 5               /* ECMAgentSimulator/Simulator.cpp:55 */
#define EO_LOOKAHEAD 10.0f   /* ECMAgentSimulator/ORCA.h:102-103 */

typedef struct { float x, y; } v2;

typedef struct { v2 n, p; } eo_constraint; /* ORCA.h:74-75 (m_N, m_PointOnLine); the slope/intercept fields are never read */

struct eo_sim {
    /* world */
    int nV, nE, nO;
    float* vert_xy; float* vert_clear; int* edge_v; float* edge_cl;
    float* obst_xy; int* obst_next; int* obst_prev; uint8_t* obst_convex;
    /* agents (Simulator.h:160-187) */
    int max_agents, num_agents, last_idx, knn_mode;
    float step;
    v2 *pos, *vel, *prefvel, *attraction, *force;
    float *radius, *speed;
    uint8_t* active;
    int* path_len; float** path_xy;
    int* free_stack; int free_top; /* Simulator.h:66-69: lowest index on top at start, LIFO reuse */
    /* neighbour structure */
    int* tree; int tree_size; int max_depth; /* KDTree.h:77-78 */
    int* sorted; int n_sorted;               /* exact mode: x-sorted active slots */
    int nn_cache[EO_K];                      /* ORCA::m_NeighborCache, ORCA.h:100 */
    /* scratch */
    int* obst_list; int obst_cap;
    eo_constraint* cons; eo_constraint* proj; int cons_cap;
    /* events */
    int* replans; int n_replans; int* destroyed; int n_destroyed;
    long long counters[8];
};

/* ---------------------------------------------------------------- primitives */
/* ECMGenerator/ECMDataTypes.h:13-60 (Vec2), :65-104 (Point), ECMGenerator/UtilityFunctions.cpp:172-321 */
static inline v2 V(float x, float y) { v2 r; r.x = x; r.y = y; return r; }
static inline v2 vadd(v2 a, v2 b) { return V(a.x + b.x, a.y + b.y); }
static inline v2 vsub(v2 a, v2 b) { return V(a.x - b.x, a.y - b.y); }
static inline v2 vmul(v2 a, float s) { return V(a.x * s, a.y * s); }
static inline v2 vdiv(v2 a, float s) { return V(a.x / s, a.y / s); }
static inline float vdot(v2 a, v2 b) { return a.x * b.x + a.y * b.y; }        /* UtilityFunctions.cpp:172-175 */
static inline float vdet(v2 a, v2 b) { return a.x * b.y - a.y * b.x; }        /* UtilityFunctions.cpp:182-191 */
static inline float vlen2(v2 a) { return a.x * a.x + a.y * a.y; }             /* ECMDataTypes.h:38-41 */
static inline float vlen(v2 a) { return sqrtf(a.x * a.x + a.y * a.y); }       /* ECMDataTypes.h:33-36 */
static inline v2 vright(v2 a) { return V(a.y, -a.x); }                        /* UtilityFunctions.cpp:213-216 */
static inline v2 vleft(v2 a) { return V(-a.y, a.x); }                         /* UtilityFunctions.cpp:223-226 */
static inline v2 vnormalized(v2 a) {                                          /* ECMDataTypes.h:45-60 */
    float l = vlen(a);
    if (l == 0.0f) return a; /* Normalize() leaves a zero vector unchanged */
    return V(a.x / l, a.y / l);
}
/* Vec2::Normalized() returns an uninitialised Vec2() for zero length (ECMDataTypes.h:54-60); we
 * return (0,0), which is what the survey's contract states (Appendix B.13). */
static inline v2 vnormalized_copy(v2 a) {
    float l = vlen(a);
    if (l == 0.0f) return V(0.0f, 0.0f);
    return V(a.x / l, a.y / l);
}
static inline int approx(v2 a, v2 b) { /* ECMDataTypes.cpp:97-100 */
    return a.x < (b.x + EO_EPSILON) && a.x > (b.x - EO_EPSILON) && a.y < (b.y + EO_EPSILON) && a.y > (b.y - EO_EPSILON);
}
static inline float sqdist_pp(v2 p1, v2 p2) { /* UtilityFunctions.cpp:34-41 */
    float dx = p2.x - p1.x, dy = p2.y - p1.y;
    return dx * dx + dy * dy;
}
static inline float sqdist_ff(float x0, float y0, float x1, float y1) { /* UtilityFunctions.cpp:43-49 */
    float dx = x0 - x1, dy = y0 - y1;
    return dx * dx + dy * dy;
}
static v2 closest_on_segment(v2 point, v2 s1, v2 s2) { /* UtilityFunctions.cpp:287-306 */
    if (approx(s1, s2)) return s1;
    v2 seg = vsub(s2, s1);
    v2 pts = vsub(point, s1);
    float tsq = sqdist_pp(s1, s2);
    float d = (pts.x * seg.x + pts.y * seg.y) / tsq;
    if (d > 1.0) d = 1.0f;
    if (d < 0.0) d = 0.0f;
    return V(s1.x + d * seg.x, s1.y + d * seg.y);
}
static v2 rotate(v2 v, float rad) { /* UtilityFunctions.cpp:233-242 */
    float cs = cosf(rad), sn = sinf(rad);
    return V(v.x * cs - v.y * sn, v.x * sn + v.y * cs);
}

static inline v2 vert(const eo_sim* s, int v) { return V(s->vert_xy[2 * v], s->vert_xy[2 * v + 1]); }
static inline v2 ecl(const eo_sim* s, int e, int k) { return V(s->edge_cl[8 * e + 2 * k], s->edge_cl[8 * e + 2 * k + 1]); }
static inline v2 obst(const eo_sim* s, int o) { return V(s->obst_xy[2 * o], s->obst_xy[2 * o + 1]); }

/* ---------------------------------------------------------------- point location */
/* MathUtility::Contains(Point, vector<Segment>), UtilityFunctions.cpp:54-86, for the closed chain q0..q3 */
static int contains4(v2 p, const v2 q[4]) {
    int inside = 0;
    for (int k = 0; k < 4; k++) {
        v2 a = q[k], b = q[(k + 1) & 3];
        if (approx(p, a)) return 0;
        if (approx(p, b)) return 0;
        if (p.y > fminf(a.y, b.y)) {
            if (p.y < fmaxf(a.y, b.y)) {
                if (p.x < fmaxf(a.x, b.x)) {
                    float xi = (p.y - a.y) * (b.x - a.x) / (b.y - a.y) + a.x;
                    if (a.x == b.x || p.x < xi) inside = !inside;
                }
            }
        }
    }
    return inside;
}
/* ECMCellCollection::PointLocationQueryLinear, ECMCellCollection.cpp:57-90; cells in construction
 * order (ECMCellCollection.cpp:17-45): cell 2e = left (L0,L1), cell 2e+1 = right (R0,R1). */
static int find_cell(const eo_sim* s, v2 p) {
    for (int c = 0; c < 2 * s->nE; c++) {
        int e = c >> 1, side = c & 1;
        v2 q[4];
        q[0] = vert(s, s->edge_v[2 * e]);     /* vertex(he[1].target) = v0 */
        q[1] = ecl(s, e, side ? 1 : 0);       /* boundary.p0 */
        q[2] = ecl(s, e, side ? 3 : 2);       /* boundary.p1 */
        q[3] = vert(s, s->edge_v[2 * e + 1]); /* vertex(he[0].target) = v1 */
        if (contains4(p, q)) return c;
    }
    return -1;
}

/* MathUtility::GetRayToLineSegmentIntersection, UtilityFunctions.cpp:323-349 */
static int ray_segment(v2 origin, v2 dir, v2 p1, v2 p2, v2* out) {
    v2 v1 = vsub(origin, p1), v2_ = vsub(p2, p1), v3 = V(-dir.y, dir.x);
    float dot = vdot(v2_, v3);
    if (fabsf(dot) < 0.000001) return 0; /* float |dot| compared against the double literal */
    float t1 = vdet(v2_, v1) / dot;
    float t2 = vdot(v1, v3) / dot;
    if (t1 >= 0.0 && (t2 >= 0.0 && t2 <= 1.0)) {
        out->x = origin.x + dir.x * t1;
        out->y = origin.y + dir.y * t1;
        return 1;
    }
    return 0;
}

/* ECM::RetractPoint, ECM.cpp:20-96 */
static int retract(const eo_sim* s, v2 loc, v2* out, int* out_edge) {
    int cell = find_cell(s, loc);
    if (cell < 0) return 0;
    int e = cell >> 1;
    *out_edge = e;
    v2 p1 = vert(s, s->edge_v[2 * e]), p2 = vert(s, s->edge_v[2 * e + 1]);
    v2 o1, o2, ray;
    /* IsLeftOfSegment, UtilityFunctions.cpp:193-196 */
    int left = (p2.x - p1.x) * (loc.y - p1.y) - (p2.y - p1.y) * (loc.x - p1.x) > 0;
    if (left) { o1 = ecl(s, e, 0); o2 = ecl(s, e, 2); }  /* he[0].closest_left, he[1].closest_right */
    else      { o1 = ecl(s, e, 1); o2 = ecl(s, e, 3); }  /* he[0].closest_right, he[1].closest_left */
    if (approx(o1, o2)) {
        v2 a = vsub(p1, o1), b = vsub(p2, o1);
        ray = vadd(a, b);
    } else {
        v2 v = vsub(o2, o1);
        ray = left ? V(v.y, -v.x) : V(-v.y, v.x);
    }
    ray = vnormalized(ray);
    return ray_segment(loc, ray, p1, p2, out);
}

/* ---------------------------------------------------------------- IRM attraction point */
/* IRMPathFollower::FindAttractionPoint, IRMPathFollower.cpp:14-115.  *out is only written where
 * the reference writes outPoint. */
static int find_attraction_point(eo_sim* s, v2 position, const float* pxy, int np, v2* out) {
    v2 R;
    int e;
    if (!retract(s, position, &R, &e)) return 0;
    v2 obstA = ecl(s, e, 0), obstB = ecl(s, e, 2); /* he[0].closest_left, he[1].closest_right */
    v2 closest = closest_on_segment(R, obstA, obstB);
    float clearance = vlen(vsub(R, closest));
    float c2 = clearance * clearance;
    v2 goal = V(pxy[2 * (np - 1)], pxy[2 * (np - 1) + 1]);
    v2 to_goal = vsub(goal, R);
    if (vlen2(to_goal) < c2) { *out = goal; return 1; }
    int success = 0;
    for (int i = 0; i < np - 1; i++) {
        v2 p1 = vsub(V(pxy[2 * i], pxy[2 * i + 1]), R);
        v2 p2 = vsub(V(pxy[2 * i + 2], pxy[2 * i + 3]), R);
        v2 ed = vsub(p2, p1);
        float el2 = vlen2(ed);
        float det = vdet(p1, p2);
        float disc = c2 * el2 - det * det;
        if (disc < EO_EPSILON) continue;
        success = 1;
        int dysign = ed.y < 0.0f ? -1 : 1;
        float sq = sqrtf(disc);
        v2 i1 = V((det * ed.y + dysign * ed.x * sq) / el2, (-det * ed.x + fabsf(ed.y) * sq) / el2);
        v2 i2 = V((det * ed.y - dysign * ed.x * sq) / el2, (-det * ed.x - fabsf(ed.y) * sq) / el2);
        v2 g1 = vadd(p1, R), g2 = vadd(p2, R), gi1 = vadd(i1, R), gi2 = vadd(i2, R);
        v2 edge = vsub(g2, g1);
        float t1 = vdot(vsub(gi1, g1), edge) / el2;
        float t2 = vdot(vsub(gi2, g1), edge) / el2;
        float maxT = -1.0f;
        if (t1 >= 0.0f && t1 <= 1.0f) { *out = gi1; maxT = t1; }
        if (t2 >= 0.0f && t2 <= 1.0f) { if (t2 > maxT) *out = gi2; }
    }
    return success;
}

/* Simulator::DestroyAgent, Simulator.cpp:202-208 */
static void destroy_agent(eo_sim* s, int idx) {
    s->num_agents--;
    s->active[idx] = 0;
    s->free_stack[s->free_top++] = idx;
}

/* Simulator::UpdateAttractionPointSystem, Simulator.cpp:538-590 */
static void update_attraction(eo_sim* s) {
    const float deleteDistanceSq = 2.0f * 2.0f, arrivalRadiusSq = 20.0f * 20.0f;
    for (int i = 0; i <= s->last_idx; i++) {
        if (!s->active[i]) continue;
        const float* pxy = s->path_xy[i];
        int np = s->path_len[i];
        v2 pos = s->pos[i];
        float gx = pxy[2 * (np - 1)], gy = pxy[2 * (np - 1) + 1];
        float d = sqdist_ff(pos.x, pos.y, gx, gy);
        if (d < arrivalRadiusSq) {
            s->attraction[i] = V(gx, gy);
            if (d < deleteDistanceSq) { destroy_agent(s, i); s->destroyed[s->n_destroyed++] = i; }
        } else {
            v2 ap = V(0.0f, 0.0f); /* `Point attractionPoint;` default-constructs to (0,0), Simulator.cpp:570 */
            if (find_attraction_point(s, pos, pxy, np, &ap)) {
                s->attraction[i] = ap;
            } else {
                s->replans[s->n_replans++] = i; /* UpdatePath(e, currentPosition, goal): host planner */
                s->counters[3]++;
            }
        }
    }
}

/* Simulator::ApplySteeringForce, Simulator.cpp:638-657 */
static void apply_steering(eo_sim* s) {
    for (int i = 0; i <= s->last_idx; i++) {
        if (!s->active[i]) continue;
        v2 d = vsub(s->attraction[i], s->pos[i]);
        d = vnormalized(d);
        s->prefvel[i] = vmul(d, s->speed[i]);
    }
}

/* ---------------------------------------------------------------- neighbours: reference KD-tree */
/* KDTree.cpp:65-72 sorts with std::sort and a strict '<' on one coordinate (KDTreeCompareX/Y,
 * KDTree.h:22-52).  The order of agents with EQUAL coordinates is implementation-defined, and
 * agents walking along axis-aligned streets do tie (observed every ~20 ticks at 1.5k agents).  To
 * stay bit-identical with the reference as compiled here we restate the published algorithm of
 * libstdc++ 13's std::sort (introsort: median-of-3 quicksort to depth 2*floor(log2 n), heapsort
 * fallback, threshold-16 final insertion sort; bits/stl_algo.h, bits/stl_heap.h). */
typedef struct { const v2* pos; int axis; } sort_ctx;
static inline int sless(const sort_ctx* c, int a, int b) {
    return c->axis ? (c->pos[a].y < c->pos[b].y) : (c->pos[a].x < c->pos[b].x);
}
static inline void iswap(int* a, int* b) { int t = *a; *a = *b; *b = t; }
static void ss_push_heap(int* first, long hole, long top, int value, const sort_ctx* c) {
    long parent = (hole - 1) / 2;
    while (hole > top && sless(c, first[parent], value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}
static void ss_adjust_heap(int* first, long hole, long len, int value, const sort_ctx* c) {
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (sless(c, first[child], first[child - 1])) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    ss_push_heap(first, hole, top, value, c);
}
static void ss_heapsort(int* first, int* last, const sort_ctx* c) { /* std::partial_sort(first, last, last) */
    long len = last - first;
    if (len >= 2) { /* make_heap */
        long parent = (len - 2) / 2;
        for (;;) {
            int value = first[parent];
            ss_adjust_heap(first, parent, len, value, c);
            if (parent == 0) break;
            parent--;
        }
    }
    while (last - first > 1) { /* sort_heap */
        --last;
        int value = *last;
        *last = *first;
        ss_adjust_heap(first, 0, last - first, value, c);
    }
}
static void ss_move_median_to_first(int* result, int* a, int* b, int* cc, const sort_ctx* c) {
    if (sless(c, *a, *b)) {
        if (sless(c, *b, *cc)) iswap(result, b);
        else if (sless(c, *a, *cc)) iswap(result, cc);
        else iswap(result, a);
    } else if (sless(c, *a, *cc)) iswap(result, a);
    else if (sless(c, *b, *cc)) iswap(result, cc);
    else iswap(result, b);
}
static int* ss_unguarded_partition(int* first, int* last, int* pivot, const sort_ctx* c) {
    for (;;) {
        while (sless(c, *first, *pivot)) ++first;
        --last;
        while (sless(c, *pivot, *last)) --last;
        if (!(first < last)) return first;
        iswap(first, last);
        ++first;
    }
}
static void ss_introsort_loop(int* first, int* last, long depth_limit, const sort_ctx* c) {
    while (last - first > 16) {
        if (depth_limit == 0) { ss_heapsort(first, last, c); return; }
        --depth_limit;
        int* mid = first + (last - first) / 2;
        ss_move_median_to_first(first, first + 1, mid, last - 1, c);
        int* cut = ss_unguarded_partition(first + 1, last, first, c);
        ss_introsort_loop(cut, last, depth_limit, c);
        last = cut;
    }
}
static void ss_unguarded_linear_insert(int* last, const sort_ctx* c) {
    int val = *last;
    int* next = last - 1;
    while (sless(c, val, *next)) { *last = *next; last = next; --next; }
    *last = val;
}
static void ss_insertion_sort(int* first, int* last, const sort_ctx* c) {
    if (first == last) return;
    for (int* i = first + 1; i != last; ++i) {
        if (sless(c, *i, *first)) {
            int val = *i;
            memmove(first + 1, first, (size_t)(i - first) * sizeof(int));
            *first = val;
        } else ss_unguarded_linear_insert(i, c);
    }
}
static void std_sort(int* first, int* last, const sort_ctx* c) {
    if (first == last) return;
    long n = last - first, lg = 0;
    while ((n >> (lg + 1)) > 0) lg++;
    ss_introsort_loop(first, last, lg * 2, c);
    if (last - first > 16) {
        ss_insertion_sort(first, first + 16, c);
        for (int* i = first + 16; i != last; ++i) ss_unguarded_linear_insert(i, c);
    } else ss_insertion_sort(first, last, c);
}
/* KDTree::ConstructRecursive, KDTree.cpp:59-83 */
static void kd_build(eo_sim* s, int index, int l, int r, int depth, int* idx) {
    if (l >= r) return;
    sort_ctx ctx = {s->pos, depth % 2};
    std_sort(idx + l, idx + r, &ctx);
    int mid = l + (r - 1 - l) / 2;
    s->tree[index] = idx[mid];
    kd_build(s, index * 2 + 1, l, mid, depth + 1, idx);
    kd_build(s, index * 2 + 2, mid + 1, r, depth + 1, idx);
}
/* KDTree::Construct, KDTree.cpp:22-57 */
static void kd_construct(eo_sim* s) {
    int size = s->num_agents;
    if (size == 0) return; /* the previous tree is kept, as in the reference */
    int* idx = (int*)malloc(sizeof(int) * (size_t)size);
    int n = 0;
    for (int i = 0; i <= s->last_idx; i++)
        if (s->active[i]) idx[n++] = i;
    s->max_depth = (int)ceil(log2((double)(size + 1)) - 1);
    int tree_size = (int)pow(2, (s->max_depth + 1)) - 1;
    if (tree_size > s->tree_size) { s->tree = (int*)realloc(s->tree, sizeof(int) * (size_t)tree_size); }
    s->tree_size = tree_size;
    for (int i = 0; i < tree_size; i++) s->tree[i] = -1;
    kd_build(s, 0, 0, size, 0, idx);
    free(idx);
}
/* KDTree::KNearestAgents_R, KDTree.cpp:98-202 - including its pruning and fill quirks */
static void kd_knn_r(const eo_sim* s, v2 target, int cur, int k, int* kFound, int depth, int* ids, float* dist) {
    if (depth > s->max_depth) return;
    if (s->tree[cur] == -1) return;
    v2 cp = s->pos[s->tree[cur]];
    v2 diff = V(cp.x - target.x, cp.y - target.y);
    float sq = diff.x * diff.x + diff.y * diff.y;
    if (*kFound < k && sq > EO_EPSILON) {
        ids[*kFound] = s->tree[cur];
        dist[*kFound] = sq;
        (*kFound)++;
        if (*kFound == k) {
            float largest = sq;
            int li = k - 1;
            for (int i = 0; i < (k - 1); i++)
                if (dist[i] > largest) { largest = dist[i]; li = i; }
            dist[0] = largest;
            ids[0] = ids[li];
            dist[li] = sq;
            ids[li] = s->tree[cur];
        }
    } else {
        *kFound = k;
        if (sq < dist[0] && sq > EO_EPSILON) {
            dist[0] = sq;
            ids[0] = s->tree[cur];
            float largest = sq;
            int li = 0;
            for (int i = 1; i < k; i++)
                if (dist[i] > largest) { largest = dist[i]; li = i; }
            dist[0] = largest;
            ids[0] = ids[li];
            dist[li] = sq;
            ids[li] = s->tree[cur];
        }
    }
    float cv = depth % 2 == 0 ? cp.x : cp.y;
    float tv = depth % 2 == 0 ? target.x : target.y;
    if (tv < cv) {
        kd_knn_r(s, target, cur * 2 + 1, k, kFound, depth + 1, ids, dist);
        float d = tv - cv, sd = d * d;
        if (sd < dist[k - 1]) kd_knn_r(s, target, cur * 2 + 2, k, kFound, depth + 1, ids, dist);
    } else {
        kd_knn_r(s, target, cur * 2 + 2, k, kFound, depth + 1, ids, dist);
        float d = tv - cv, sd = d * d;
        if (sd < dist[k - 1]) kd_knn_r(s, target, cur * 2 + 1, k, kFound, depth + 1, ids, dist);
    }
}

/* ---------------------------------------------------------------- neighbours: exact kNN spec */
/* SURVEY.md §8c "exact-knn": candidates = active at Construct; sqDist = fl(fl(dx*dx)+fl(dy*dy));
 * keep sqDist > 1e-4f; k smallest by (sqDist, slot) ascending; unused slots = -1. */
static int cmp_x(const void* a, const void* b, void* c) {
    const v2* pos = (const v2*)c;
    int ia = *(const int*)a, ib = *(const int*)b;
    if (pos[ia].x < pos[ib].x) return -1;
    if (pos[ia].x > pos[ib].x) return 1;
    return (ia > ib) - (ia < ib);
}
static void exact_construct(eo_sim* s) {
    int n = 0;
    for (int i = 0; i <= s->last_idx; i++)
        if (s->active[i]) s->sorted[n++] = i;
    s->n_sorted = n;
    qsort_r(s->sorted, (size_t)n, sizeof(int), cmp_x, s->pos);
}
static inline void exact_consider(const eo_sim* s, v2 t, int j, int k, int* found, int* ids, float* best) {
    float dx = s->pos[j].x - t.x, dy = s->pos[j].y - t.y;
    float mx = dx * dx, my = dy * dy;
    float d = mx + my;
    if (!(d > EO_EPSILON)) return;
    int p = *found < k ? *found : k;
    while (p > 0 && (d < best[p - 1] || (d == best[p - 1] && j < ids[p - 1]))) p--;
    if (p >= k) return;
    for (int q = (*found < k ? *found : k - 1); q > p; q--) { best[q] = best[q - 1]; ids[q] = ids[q - 1]; }
    best[p] = d;
    ids[p] = j;
    if (*found < k) (*found)++;
}
static void exact_knn(const eo_sim* s, int agent, int k, int* ids, int* count) {
    v2 t = s->pos[agent];
    float best[EO_K];
    int found = 0, n = s->n_sorted;
    for (int i = 0; i < k; i++) { best[i] = EO_MAX_FLOAT; ids[i] = -1; }
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) / 2;
        if (s->pos[s->sorted[mid]].x < t.x) lo = mid + 1; else hi = mid;
    }
    int r = lo, l = lo - 1, goR = 1, goL = 1;
    while (goR || goL) {
        if (goR) {
            if (r >= n) goR = 0;
            else {
                float dx = s->pos[s->sorted[r]].x - t.x;
                if (found == k && dx * dx > best[k - 1]) goR = 0;
                else exact_consider(s, t, s->sorted[r++], k, &found, ids, best);
            }
        }
        if (goL) {
            if (l < 0) goL = 0;
            else {
                float dx = s->pos[s->sorted[l]].x - t.x;
                if (found == k && dx * dx > best[k - 1]) goL = 0;
                else exact_consider(s, t, s->sorted[l--], k, &found, ids, best);
            }
        }
    }
    *count = found;
}

/* Simulator::FindNNearestNeighbors -> KDTree::KNearestAgents, Simulator.cpp:211-227, KDTree.cpp:85-96 */
static void find_neighbors(const eo_sim* s, int agent, int* ids, int* count) {
    if (s->knn_mode == EO_KNN_EXACT) { exact_knn(s, agent, EO_K, ids, count); return; }
    float dist[EO_K];
    for (int i = 0; i < EO_K; i++) dist[i] = EO_MAX_FLOAT;
    *count = 0;
    if (s->tree_size > 0) kd_knn_r(s, s->pos[agent], 0, EO_K, count, 0, ids, dist);
}

/* ---------------------------------------------------------------- obstacles */
/* Simulator::FindNearestObstacles, Simulator.cpp:259-292 */
static int find_obstacles(const eo_sim* s, int agent, float range2, int* out, int cap) {
    v2 a = s->pos[agent];
    int n = 0;
    for (int o = 0; o < s->nO; o++) {
        v2 p = obst(s, o), q = obst(s, s->obst_next[o]);
        /* LineLeftDistance(v1, v2, p1) = Determinant(v1 - p1, v2 - v1), UtilityFunctions.cpp:49-52 */
        float sl = vdet(vsub(p, a), vsub(q, p));
        float sq = powf(sl, 2.0f) / sqdist_pp(p, q);
        if (sq < range2) {
            if (sl < 0.0f) {
                v2 c = closest_on_segment(a, p, q);
                if (sqdist_pp(c, a) < range2) {
                    if (n < cap) out[n] = o;
                    n++;
                }
            }
        }
    }
    return n;
}

/* ---------------------------------------------------------------- ORCA constraints */
static inline void cinit(eo_constraint* c, v2 point, v2 normal) { c->n = normal; c->p = point; } /* ORCA.h:29-47 */

/* ORCA::GenerateConstraints, ORCA.cpp:60-424.  Returns the total count; *n_obst = obstacle part. */
static int generate_constraints(const eo_sim* s, int entity, int n_nb, const int* nb, int n_on, const int* on,
                                float stepSize, int* n_obst, eo_constraint* out) {
    v2 position = s->pos[entity];
    float clearance = s->radius[entity];
    v2 velocity = s->vel[entity];
    int nc = 0;
    for (int i = 0; i < n_on; i++) {
        int oL = on[i];
        int oR = s->obst_next[on[i]];
        v2 rp1 = vsub(obst(s, oL), position);
        v2 rp2 = vsub(obst(s, oR), position);
        v2 segDir = vsub(obst(s, oR), obst(s, oL));
        const float sp = vdot(vmul(rp1, -1.0f), segDir) / vlen2(segDir);
        const float distSqLine = vlen2(vsub(vmul(rp1, -1.0f), vmul(segDir, sp)));
        const float distSq1 = vlen2(rp1);
        const float distSq2 = vlen2(rp2);
        segDir = vnormalized(segDir);
        const float radiusSq = clearance * clearance;

        if (sp < 0.0f && distSq1 <= radiusSq) { /* ORCA.cpp:92-105 */
            if (s->obst_convex[oL]) cinit(&out[nc++], V(0.0f, 0.0f), vnormalized_copy(vmul(rp1, -1.0f)));
            continue;
        } else if (sp > 1.0f && distSq2 <= radiusSq) { /* ORCA.cpp:108-121 */
            v2 rnd = vsub(obst(s, s->obst_next[oR]), obst(s, oR));
            rnd = vnormalized(rnd);
            if (s->obst_convex[oR] && vdet(rp2, rnd) >= 0.0f) cinit(&out[nc++], V(0.0f, 0.0f), vnormalized_copy(vmul(rp2, -1.0f)));
            continue;
        } else if (sp >= 0.0f && sp < 1.0f && distSqLine <= radiusSq) { /* ORCA.cpp:124-135 */
            cinit(&out[nc++], V(0.0f, 0.0f), vright(segDir));
            continue;
        }

        v2 leftLeg, rightLeg;
        if (sp < 0.0f && distSqLine <= radiusSq) { /* ORCA.cpp:146-169 */
            if (!s->obst_convex[oL]) continue;
            oR = oL;
            const float leg1 = sqrtf(distSq1 - radiusSq);
            leftLeg = vdiv(V(rp1.x * leg1 - rp1.y * clearance, rp1.x * clearance + rp1.y * leg1), distSq1);
            rightLeg = vdiv(V(rp1.x * leg1 + rp1.y * clearance, -rp1.x * clearance + rp1.y * leg1), distSq1);
        } else if (sp > 1.0f && distSqLine <= radiusSq) { /* ORCA.cpp:171-183 */
            if (!s->obst_convex[oR]) continue;
            oL = oR;
            const float leg2 = sqrtf(distSq2 - radiusSq);
            leftLeg = vdiv(V(rp2.x * leg2 - rp2.y * clearance, rp2.x * clearance + rp2.y * leg2), distSq2);
            rightLeg = vdiv(V(rp2.x * leg2 + rp2.y * clearance, -rp2.x * clearance + rp2.y * leg2), distSq2);
        } else { /* ORCA.cpp:186-212 */
            if (s->obst_convex[oL]) {
                const float leg1 = sqrtf(distSq1 - radiusSq);
                leftLeg = vdiv(V(rp1.x * leg1 - rp1.y * clearance, rp1.x * clearance + rp1.y * leg1), distSq1);
            } else {
                leftLeg = vmul(segDir, -1.0f);
            }
            if (s->obst_convex[oR]) {
                const float leg2 = sqrtf(distSq2 - radiusSq);
                rightLeg = vdiv(V(rp2.x * leg2 + rp2.y * clearance, -rp2.x * clearance + rp2.y * leg2), distSq2);
            } else {
                rightLeg = segDir;
            }
        }

        /* foreign legs, ORCA.cpp:218-239 */
        int leftNeighbor = s->obst_prev[oL];
        int isLeftLegForeign = 0, isRightLegForeign = 0;
        v2 lnd = vsub(obst(s, oL), obst(s, leftNeighbor));
        lnd = vnormalized(lnd);
        if (s->obst_convex[oL] && vdet(leftLeg, vmul(lnd, -1.0f)) >= 0.0f) { leftLeg = vmul(lnd, -1.0f); isLeftLegForeign = 1; }
        v2 rnd = vsub(obst(s, s->obst_next[oR]), obst(s, oR));
        rnd = vnormalized(rnd);
        if (s->obst_convex[oR] && vdet(rightLeg, rnd) <= 0.0f) { rightLeg = rnd; isRightLegForeign = 1; }

        float recip = 1.0f / EO_LOOKAHEAD; /* ORCA.cpp:241 */
        const v2 leftCutoff = vmul(vsub(obst(s, oL), position), recip);
        const v2 rightCutoff = vmul(vsub(obst(s, oR), position), recip);
        const v2 cutoffVec = vsub(rightCutoff, leftCutoff);
        const float t = (oL == oR ? 0.5f : vdot(vsub(velocity, leftCutoff), cutoffVec) / vlen2(cutoffVec));
        const float tLeft = vdot(vsub(velocity, leftCutoff), leftLeg);
        const float tRight = vdot(vsub(velocity, rightCutoff), rightLeg);

        if ((t < 0.0f && tLeft < 0.0f) || (oL == oR && tLeft < 0.0f && tRight < 0.0f)) { /* ORCA.cpp:259-268 */
            v2 unitW = vnormalized_copy(vsub(velocity, leftCutoff));
            v2 pc = vadd(leftCutoff, vmul(vmul(unitW, recip), clearance));
            cinit(&out[nc++], pc, unitW);
            continue;
        } else if (t > 1.0f && tRight < 0.0f) { /* ORCA.cpp:270-280 */
            v2 unitW = vnormalized_copy(vsub(velocity, rightCutoff));
            v2 pc = vadd(rightCutoff, vmul(vmul(unitW, recip), clearance));
            cinit(&out[nc++], pc, unitW);
            continue;
        }
        /* ORCA.cpp:284-286 */
        const float distSqCutoff = ((t < 0.0f || t > 1.0f || oL == oR) ? INFINITY : vlen2(vsub(velocity, vadd(leftCutoff, vmul(cutoffVec, t)))));
        const float distSqLeft = ((tLeft < 0.0f) ? INFINITY : vlen2(vsub(velocity, vadd(leftCutoff, vmul(leftLeg, tLeft)))));
        const float distSqRight = ((tRight < 0.0f) ? INFINITY : vlen2(vsub(velocity, vadd(rightCutoff, vmul(rightLeg, tRight)))));

        if (distSqCutoff <= distSqLeft && distSqCutoff <= distSqRight) { /* ORCA.cpp:289-301 */
            v2 normal = vleft(vmul(segDir, -1.0f));
            v2 pc = vadd(leftCutoff, vmul(vmul(normal, recip), clearance));
            cinit(&out[nc++], pc, normal);
            continue;
        } else if (distSqLeft <= distSqRight) { /* ORCA.cpp:303-317 */
            if (isLeftLegForeign) continue;
            v2 normal = vleft(leftLeg);
            v2 pc = vadd(leftCutoff, vmul(vmul(normal, clearance), recip));
            cinit(&out[nc++], pc, normal);
            continue;
        } else { /* ORCA.cpp:319-332 */
            if (isRightLegForeign) continue;
            v2 normal = vright(rightLeg);
            v2 pc = vadd(rightCutoff, vmul(vmul(normal, clearance), recip));
            cinit(&out[nc++], pc, normal);
            continue;
        }
    }
    *n_obst = nc;

    /* agent constraints, ORCA.cpp:339-423 */
    for (int i = 0; i < n_nb; i++) {
        int nbr = nb[i];
        v2 np_ = s->pos[nbr], nv = s->vel[nbr];
        float ncl = s->radius[nbr];
        v2 VOPos = V((np_.x - position.x) / EO_LOOKAHEAD, (np_.y - position.y) / EO_LOOKAHEAD);
        float VOPosLength = vlen(VOPos);
        float combinedRadius = ncl + clearance;
        float VORadius = combinedRadius / EO_LOOKAHEAD;
        v2 relVel = V(velocity.x - nv.x, velocity.y - nv.y);
        v2 relPos = V(np_.x - position.x, np_.y - position.y);
        float relPosLength = vlen(relPos);
        if (relPosLength < combinedRadius) { /* ORCA.cpp:356-372 */
            v2 w = vsub(relVel, vdiv(relPos, stepSize));
            float wLength = vlen(w);
            v2 unitW = vdiv(w, wLength);
            v2 U = vmul(unitW, (combinedRadius / stepSize - wLength));
            cinit(&out[nc++], vadd(velocity, vmul(U, 0.5f)), unitW);
            continue;
        }
        float tanAngleFactor = VORadius / VOPosLength;
        float tanHalfAngle = atanf(tanAngleFactor);
        v2 VOLeftLeg = rotate(VOPos, tanHalfAngle);
        v2 VORightLeg = rotate(VOPos, -tanHalfAngle);
        float sqDistFromCircleCentre = sqdist_pp(VOPos, relVel);
        v2 base = vsub(VOLeftLeg, VOPos), chk = vsub(relVel, VOPos);
        int liesBelow = base.x * chk.y - base.y * chk.x > 0; /* IsLeftOfVector, UtilityFunctions.cpp:198-201 */
        if (liesBelow) { /* ORCA.cpp:385-397 */
            float distToEdge = VORadius - sqrtf(sqDistFromCircleCentre);
            v2 lineNormal = vnormalized(vsub(relVel, VOPos));
            v2 pointOnLine = vadd(velocity, vmul(vmul(lineNormal, distToEdge), 0.5f));
            cinit(&out[nc++], pointOnLine, lineNormal);
        } else { /* ORCA.cpp:399-421 */
            v2 leftPerp = vdiv(vleft(VOPos), VOPosLength);
            int closerToLeft = vdot(leftPerp, relVel) >= 0;
            if (closerToLeft) {
                v2 ln = vnormalized_copy(VOLeftLeg);
                float l = vdot(relVel, ln); /* GetClosestPointOnLineThroughOrigin, UtilityFunctions.cpp:316-320 */
                v2 U = vsub(vmul(ln, l), relVel);
                cinit(&out[nc++], vadd(velocity, vmul(U, 0.5f)), vleft(ln));
            } else {
                v2 rn = vnormalized_copy(VORightLeg);
                float l = vdot(relVel, rn);
                v2 U = vsub(vmul(rn, l), relVel);
                cinit(&out[nc++], vadd(velocity, vmul(U, 0.5f)), vright(rn));
            }
        }
    }
    return nc;
}

/* Constraint::Contains (methodB), ORCA.h:49-60 */
static inline int ccontains(const eo_constraint* c, v2 p) { return vdet(vright(c->n), vsub(c->p, p)) <= 0.0f; }

/* ORCA::RandomizedLP, ORCA.cpp:428-587 */
static int randomized_lp(eo_sim* s, const eo_constraint* cs, int n, v2 opt, float maxSpeed, int useDirOpt, v2* outV) {
    s->counters[1]++;
    if (useDirOpt) *outV = vmul(opt, maxSpeed);
    else if (vlen(opt) > maxSpeed) *outV = vmul(vnormalized_copy(opt), maxSpeed);
    else *outV = opt;
    if (n == 0) return n;
    for (int i = 0; i < n; i++) {
        const eo_constraint* h = &cs[i];
        if (ccontains(h, *outV)) continue;
        v2 dir = vright(h->n);
        float dpd = vdot(dir, h->p);
        float disc = dpd * dpd + maxSpeed * maxSpeed - vdot(h->p, h->p);
        if (disc <= 0.0f) return i;
        if (!(disc > 0.0f)) continue; /* NaN: `if (d <= 0) return i; else if (d > 0) {...}` does neither (ORCA.cpp:499-507) */
        float dsq = sqrtf(disc);
        float left = -dpd - dsq;
        float right = -dpd + dsq;
        for (int j = 0; j < i; j++) {
            const eo_constraint* hj = &cs[j];
            float den = vdet(vright(h->n), vright(hj->n));
            float num = vdet(vright(hj->n), vsub(h->p, hj->p));
            if (fabsf(den) <= EO_EPSILON) {
                if (num < 0.0f) return i;
                else continue;
            }
            const float t = num / den;
            if (den >= 0.0f) right = (t < right) ? t : right; /* std::min(right, t) */
            else left = (left < t) ? t : left;               /* std::max(left, t) */
            if (left > right) return i;
        }
        if (useDirOpt) {
            if (vdot(opt, vright(h->n)) > 0) *outV = vadd(h->p, vmul(vright(h->n), right));
            else *outV = vadd(h->p, vmul(vright(h->n), left));
        } else {
            float t = vdot(vright(h->n), vsub(opt, h->p));
            if (t < left) *outV = vadd(h->p, vmul(vright(h->n), left));
            else if (t > right) *outV = vadd(h->p, vmul(vright(h->n), right));
            else *outV = vadd(h->p, vmul(vright(h->n), t));
        }
    }
    return n;
}

/* ORCA::RandomizedLP3D, ORCA.cpp:592-669 */
static void randomized_lp3d(eo_sim* s, int nObst, const eo_constraint* cs, int total, float maxSpeed, int failed, v2* outV) {
    s->counters[0]++;
    float maxPen = 0.0f;
    for (int i = failed; i < total; i++) {
        const eo_constraint* ci = &cs[i];
        v2 dir = vright(ci->n);
        if (vdet(dir, vsub(ci->p, *outV)) <= maxPen) continue;
        int np = 0;
        for (int k = 0; k < nObst; k++) s->proj[np++] = cs[k];
        for (int j = nObst; j < i; j++) {
            const eo_constraint* cj = &cs[j];
            float det = vdet(vright(ci->n), vright(cj->n));
            v2 pt;
            if (fabsf(det) <= EO_EPSILON) {
                if (vdot(ci->n, cj->n) > 0) continue;
                pt = vmul(vadd(ci->p, cj->p), 0.5f);
            } else {
                float t = vdet(vright(cj->n), vsub(ci->p, cj->p)) / det;
                pt = vadd(ci->p, vmul(vright(ci->n), t));
            }
            cinit(&s->proj[np++], pt, vnormalized_copy(vsub(cj->n, ci->n)));
        }
        const v2 temp = *outV;
        if (randomized_lp(s, s->proj, np, ci->n, maxSpeed, 1, outV) < np) *outV = temp;
        maxPen = vdet(dir, vsub(ci->p, *outV));
    }
}

/* ORCA::GetVelocity given the neighbour list, ORCA.cpp:23-56 */
static v2 orca_velocity(eo_sim* s, int entity, int n_nb, const int* nb) {
    float maxSpeed = s->speed[entity]; /* Simulator.cpp:671: m_PreferredSpeed[i] */
    float range = EO_LOOKAHEAD * maxSpeed + s->radius[entity];
    int n_on = find_obstacles(s, entity, range * range, s->obst_list, s->obst_cap);
    if (n_on > s->obst_cap) { /* grow and redo: the reference list is unbounded */
        s->obst_cap = n_on * 2;
        s->obst_list = (int*)realloc(s->obst_list, sizeof(int) * (size_t)s->obst_cap);
        n_on = find_obstacles(s, entity, range * range, s->obst_list, s->obst_cap);
    }
    if (n_on > s->counters[2]) s->counters[2] = n_on;
    for (int i = 0; i < n_on; i++) { /* test coverage statistics only: what kind of geometry reached GenerateConstraints */
        const int oL = s->obst_list[i], oR = s->obst_next[oL];
        if (!s->obst_convex[oL] || !s->obst_convex[oR]) s->counters[6]++;
        const v2 a = obst(s, oL), b = obst(s, oR);
        if (a.x != b.x && a.y != b.y) s->counters[7]++;
    }
    if (n_on + EO_K > s->cons_cap) {
        s->cons_cap = (n_on + EO_K) * 2;
        s->cons = (eo_constraint*)realloc(s->cons, sizeof(eo_constraint) * (size_t)s->cons_cap);
        s->proj = (eo_constraint*)realloc(s->proj, sizeof(eo_constraint) * (size_t)s->cons_cap);
    }
    int nObst = 0;
    int n = generate_constraints(s, entity, n_nb, nb, n_on, s->obst_list, s->step, &nObst, s->cons);
    v2 out = V(0.0f, 0.0f);
    int failed = randomized_lp(s, s->cons, n, s->prefvel[entity], maxSpeed, 0, &out);
    if (failed < n) randomized_lp3d(s, nObst, s->cons, n, maxSpeed, failed, &out);
    return out;
}

/* Simulator::ApplyObstacleAvoidanceForce, Simulator.cpp:659-686 */
static void apply_orca(eo_sim* s) {
    for (int i = 0; i <= s->last_idx; i++) {
        if (!s->active[i]) continue;
        int count = 0;
        find_neighbors(s, i, s->nn_cache, &count);
        v2 out = orca_velocity(s, i, count, s->nn_cache);
        s->force[i] = V(out.x - s->vel[i].x, out.y - s->vel[i].y);
        s->counters[5]++;
    }
}

/* Simulator::UpdateVelocitySystem, Simulator.cpp:619-635 */
static void update_velocity(eo_sim* s) {
    const float mass = 0.8f;
    const float massRecip = 1.0f / mass;
    for (int i = 0; i <= s->last_idx; i++) {
        if (!s->active[i]) continue;
        s->vel[i].x += s->force[i].x * massRecip * s->step;
        s->vel[i].y += s->force[i].y * massRecip * s->step;
    }
}
/* Simulator::UpdatePositionSystem, Simulator.cpp:592-606 */
static void update_position(eo_sim* s) {
    for (int i = 0; i <= s->last_idx; i++) {
        if (!s->active[i]) continue;
        s->pos[i].x += (s->vel[i].x * s->step);
        s->pos[i].y += (s->vel[i].y * s->step);
    }
}
/* Simulator::UpdateMaxAgentIndex, Simulator.cpp:481-492 */
static void update_max_index(eo_sim* s) {
    int empty = 0;
    for (int i = s->last_idx; i >= 0; i--) {
        if (s->active[i]) break;
        empty++;
    }
    s->last_idx -= empty;
}

/* ---------------------------------------------------------------- public API */
static void* dup_mem(const void* p, size_t n) {
    void* r = malloc(n ? n : 1);
    if (n) memcpy(r, p, n);
    return r;
}

eo_sim* eo_create(int nV, const float* vert_xy, const float* vert_clear, int nE, const int* edge_v,
                  const float* edge_cl, int nO, const float* obst_xy, const int* obst_next, const int* obst_prev,
                  const uint8_t* obst_convex, int max_agents, float step, int knn_mode) {
    eo_sim* s = (eo_sim*)calloc(1, sizeof(eo_sim));
    s->nV = nV; s->nE = nE; s->nO = nO;
    s->vert_xy = (float*)dup_mem(vert_xy, sizeof(float) * 2 * (size_t)nV);
    s->vert_clear = (float*)dup_mem(vert_clear, sizeof(float) * (size_t)nV);
    s->edge_v = (int*)dup_mem(edge_v, sizeof(int) * 2 * (size_t)nE);
    s->edge_cl = (float*)dup_mem(edge_cl, sizeof(float) * 8 * (size_t)nE);
    s->obst_xy = (float*)dup_mem(obst_xy, sizeof(float) * 2 * (size_t)nO);
    s->obst_next = (int*)dup_mem(obst_next, sizeof(int) * (size_t)nO);
    s->obst_prev = (int*)dup_mem(obst_prev, sizeof(int) * (size_t)nO);
    s->obst_convex = (uint8_t*)dup_mem(obst_convex, (size_t)nO);
    s->max_agents = max_agents; s->step = step; s->knn_mode = knn_mode;
    s->last_idx = -1; /* Simulator.cpp:23 */
    size_t n = (size_t)max_agents;
    s->pos = (v2*)calloc(n, sizeof(v2)); s->vel = (v2*)calloc(n, sizeof(v2)); s->prefvel = (v2*)calloc(n, sizeof(v2));
    s->attraction = (v2*)calloc(n, sizeof(v2)); s->force = (v2*)calloc(n, sizeof(v2));
    s->radius = (float*)calloc(n, sizeof(float)); s->speed = (float*)calloc(n, sizeof(float));
    s->active = (uint8_t*)calloc(n, 1);
    s->path_len = (int*)calloc(n, sizeof(int)); s->path_xy = (float**)calloc(n, sizeof(float*));
    s->free_stack = (int*)malloc(sizeof(int) * n);
    for (int i = max_agents - 1; i >= 0; i--) s->free_stack[s->free_top++] = i; /* Simulator.h:66-69 */
    s->sorted = (int*)malloc(sizeof(int) * n);
    s->replans = (int*)malloc(sizeof(int) * n); s->destroyed = (int*)malloc(sizeof(int) * n);
    s->obst_cap = 64; s->obst_list = (int*)malloc(sizeof(int) * 64);
    s->cons_cap = 64 + EO_K;
    s->cons = (eo_constraint*)malloc(sizeof(eo_constraint) * (size_t)s->cons_cap);
    s->proj = (eo_constraint*)malloc(sizeof(eo_constraint) * (size_t)s->cons_cap);
    return s;
}

void eo_destroy(eo_sim* s) {
    if (!s) return;
    for (int i = 0; i < s->max_agents; i++) free(s->path_xy[i]);
    free(s->vert_xy); free(s->vert_clear); free(s->edge_v); free(s->edge_cl);
    free(s->obst_xy); free(s->obst_next); free(s->obst_prev); free(s->obst_convex);
    free(s->pos); free(s->vel); free(s->prefvel); free(s->attraction); free(s->force);
    free(s->radius); free(s->speed); free(s->active); free(s->path_len); free(s->path_xy);
    free(s->free_stack); free(s->tree); free(s->sorted); free(s->replans); free(s->destroyed);
    free(s->obst_list); free(s->cons); free(s->proj);
    free(s);
}

void eo_set_path(eo_sim* s, int slot, const float* xy, int n) { /* Simulator::UpdatePath, Simulator.cpp:97-124 */
    free(s->path_xy[slot]);
    s->path_xy[slot] = (float*)dup_mem(xy, sizeof(float) * 2 * (size_t)n);
    s->path_len[slot] = n;
}

int eo_bulk_load(eo_sim* s, int n, const float* pos_xy, const float* radius, const float* speed, const int* path_off,
                 const float* path_xy, int* out_slots) {
    int loaded = 0;
    for (int i = 0; i < n; i++) {
        if (out_slots) out_slots[i] = -1;
        if (s->free_top == 0) break;
        int np = path_off[i + 1] - path_off[i];
        if (np < 2) continue;
        s->num_agents++;
        int idx = s->free_stack[--s->free_top];
        s->last_idx = s->last_idx < idx ? idx : s->last_idx;
        s->pos[idx] = V(pos_xy[2 * i], pos_xy[2 * i + 1]);
        s->radius[idx] = radius[i];
        s->speed[idx] = speed[i];
        s->active[idx] = 1;
        eo_set_path(s, idx, path_xy + 2 * path_off[i], np);
        s->prefvel[idx] = s->vel[idx] = s->force[idx] = s->attraction[idx] = V(0.0f, 0.0f);
        if (out_slots) out_slots[i] = idx;
        loaded++;
    }
    return loaded;
}

void eo_set_kinematics(eo_sim* s, int slot, float x, float y, float vx, float vy) { s->pos[slot] = V(x, y); s->vel[slot] = V(vx, vy); }
void eo_set_attraction(eo_sim* s, int slot, float x, float y) { s->attraction[slot] = V(x, y); }
void eo_destroy_agent(eo_sim* s, int slot) { destroy_agent(s, slot); }

void eo_step(eo_sim* s) { /* Simulator::Update, Simulator.cpp:314-323 */
    s->n_replans = 0;
    s->n_destroyed = 0;
    update_max_index(s);
    if (s->knn_mode == EO_KNN_EXACT) exact_construct(s); else kd_construct(s);
    update_attraction(s); /* UpdateForceSystem, Simulator.cpp:611-617 */
    apply_steering(s);
    apply_orca(s);
    update_velocity(s);
    update_position(s);
}

int eo_num_replans(const eo_sim* s) { return s->n_replans; }
const int* eo_replans(const eo_sim* s) { return s->replans; }
int eo_num_destroyed(const eo_sim* s) { return s->n_destroyed; }
const int* eo_destroyed(const eo_sim* s) { return s->destroyed; }
int eo_num_agents(const eo_sim* s) { return s->num_agents; }
int eo_last_index(const eo_sim* s) { return s->last_idx; }
const long long* eo_counters(const eo_sim* s) { return s->counters; }

void eo_get_state(const eo_sim* s, int count, float* pos, float* vel, float* prefvel, float* attraction, float* force,
                  uint8_t* active) {
    size_t b = sizeof(v2) * (size_t)count;
    if (pos) memcpy(pos, s->pos, b);
    if (vel) memcpy(vel, s->vel, b);
    if (prefvel) memcpy(prefvel, s->prefvel, b);
    if (attraction) memcpy(attraction, s->attraction, b);
    if (force) memcpy(force, s->force, b);
    if (active) memcpy(active, s->active, (size_t)count);
}

void eo_query_cells(const eo_sim* s, int n, const float* xy, int* out_cell) {
    for (int i = 0; i < n; i++) out_cell[i] = find_cell(s, V(xy[2 * i], xy[2 * i + 1]));
}

void eo_retract(const eo_sim* s, int n, const float* xy, uint8_t* ok, float* out_xy, int* out_edge) {
    for (int i = 0; i < n; i++) {
        v2 r = V(0.0f, 0.0f);
        int e = -1;
        ok[i] = (uint8_t)retract(s, V(xy[2 * i], xy[2 * i + 1]), &r, &e);
        out_xy[2 * i] = r.x; out_xy[2 * i + 1] = r.y;
        out_edge[i] = e;
    }
}

void eo_query_neighbors(eo_sim* s, int count, int* out_ids, int* out_counts) {
    if (s->knn_mode == EO_KNN_EXACT) exact_construct(s); else kd_construct(s);
    int cache[EO_K] = {0, 0, 0, 0, 0};
    for (int i = 0; i < count; i++) {
        out_counts[i] = -1;
        if (i > s->last_idx || !s->active[i]) continue;
        int c = 0;
        find_neighbors(s, i, cache, &c);
        out_counts[i] = c;
        for (int k = 0; k < EO_K; k++) out_ids[EO_K * i + k] = cache[k];
    }
}

int eo_query_obstacles(const eo_sim* s, int slot, int* out_ids, int cap) {
    float range = EO_LOOKAHEAD * s->speed[slot] + s->radius[slot]; /* ORCA.cpp:27 */
    return find_obstacles(s, slot, range * range, out_ids, cap);
}

void eo_orca_velocity(const eo_sim* s, int slot, int n_neighbors, const int* neighbors, float* out_v) {
    v2 v = orca_velocity((eo_sim*)s, slot, n_neighbors, neighbors);
    out_v[0] = v.x; out_v[1] = v.y;
}

/* test hook: the std::sort restatement on its own (tests/test_oracle_vs_reference.py) */
void eo_test_std_sort(int* first, int n, const float* pos_xy, int axis) {
    sort_ctx c = {(const v2*)pos_xy, axis};
    std_sort(first, first + n, &c);
}
