#include "planner.h"

#include <algorithm>
#include <cmath>
#include <queue>
#include <thread>

// Float semantics: this file is compiled with -ffp-contract=off and no -march so that every
// expression rounds like the reference's.

namespace ecmb200 {

namespace {

constexpr float kEps = 0.0001f;  // Configuration.h:14
constexpr float kMaxFloat = 3.402823466e+38f;

inline P2f P(float x, float y) { return P2f{x, y}; }
inline P2f operator-(P2f a, P2f b) { return P(a.x - b.x, a.y - b.y); }
inline P2f operator+(P2f a, P2f b) { return P(a.x + b.x, a.y + b.y); }
inline P2f operator*(P2f a, float s) { return P(a.x * s, a.y * s); }
inline P2f operator/(P2f a, float s) { return P(a.x / s, a.y / s); }
inline bool Approx(P2f a, P2f b) {  // Point::Approximate / operator== (ECMDataTypes.cpp:97-100, ECMDataTypes.h:99-103)
    return a.x < (b.x + kEps) && a.x > (b.x - kEps) && a.y < (b.y + kEps) && a.y > (b.y - kEps);
}
inline float Length(P2f a) { return std::sqrt(a.x * a.x + a.y * a.y); }
inline P2f Normalized(P2f a) {  // Vec2::Normalize (ECMDataTypes.h:45-52)
    float l = Length(a);
    if (l == 0.0f) return a;
    return P(a.x / l, a.y / l);
}
inline float Distance(P2f p1, P2f p2) {  // MathUtility::Distance (UtilityFunctions.cpp:17-24)
    float dx = p2.x - p1.x, dy = p2.y - p1.y;
    return std::sqrt(dx * dx + dy * dy);
}
inline bool IsLeftOfSegment(P2f s0, P2f s1, P2f p) {  // UtilityFunctions.cpp:193-196
    return (s1.x - s0.x) * (p.y - s0.y) - (s1.y - s0.y) * (p.x - s0.x) > 0;
}
inline float TriangleArea(P2f p1, P2f p2, P2f p3) {  // UtilityFunctions.cpp:203-210
    float ax = p2.x - p1.x, ay = p2.y - p1.y, bx = p3.x - p1.x, by = p3.y - p1.y;
    return bx * ay - ax * by;
}

inline P2f Vert(const FlatWorld& w, int v) { return P(w.ecm.vert_xy[2 * v], w.ecm.vert_xy[2 * v + 1]); }
inline P2f Cl(const FlatWorld& w, int e, int k) { return P(w.ecm.edge_cl[8 * e + 2 * k], w.ecm.edge_cl[8 * e + 2 * k + 1]); }
// half-edge accessors (ECM.h:20-26, :69); see flat_world.h for the L0 R0 L1 R1 mapping
inline int HeTarget(const FlatWorld& w, int he) { return (he & 1) ? w.ecm.edge_v[2 * (he >> 1)] : w.ecm.edge_v[2 * (he >> 1) + 1]; }
inline int HeSource(const FlatWorld& w, int he) { return (he & 1) ? w.ecm.edge_v[2 * (he >> 1) + 1] : w.ecm.edge_v[2 * (he >> 1)]; }
inline P2f HeClosestLeft(const FlatWorld& w, int he) { return Cl(w, he >> 1, (he & 1) ? 3 : 0); }
inline P2f HeClosestRight(const FlatWorld& w, int he) { return Cl(w, he >> 1, (he & 1) ? 2 : 1); }

// MathUtility::Contains(Point, 4 segments) (UtilityFunctions.cpp:54-86)
bool Contains4(P2f p, const P2f q[4]) {
    bool inside = false;
    for (int k = 0; k < 4; k++) {
        P2f a = q[k], b = q[(k + 1) & 3];
        if (Approx(p, a)) return false;
        if (Approx(p, b)) return false;
        if (p.y > std::fmin(a.y, b.y) && p.y < std::fmax(a.y, b.y) && p.x < std::fmax(a.x, b.x)) {
            float xi = (p.y - a.y) * (b.x - a.x) / (b.y - a.y) + a.x;
            if (a.x == b.x || p.x < xi) inside = !inside;
        }
    }
    return inside;
}
bool CellContains(const FlatWorld& w, int c, P2f p) {
    int e = c >> 1, side = c & 1;
    P2f q[4] = {Vert(w, w.ecm.edge_v[2 * e]), Cl(w, e, side), Cl(w, e, 2 + side), Vert(w, w.ecm.edge_v[2 * e + 1])};
    return Contains4(p, q);
}

struct Seg {
    P2f p0, p1;
};

}  // namespace

// ------------------------------------------------------------------------------------------------
void CellLocator::Build(const FlatWorld& w, float bin) {
    const double W = (double)w.bbox[2] - w.bbox[0], H = (double)w.bbox[3] - w.bbox[1];
    const int nE = w.ecm.num_edges();
    double b = bin > 0 ? bin : 0.4 * std::sqrt(W * H / std::max(1, nE));
    while ((W / b + 3) * (H / b + 3) > 16.0e6) b *= 1.5;
    bin_ = (float)b;
    x0_ = (float)(w.bbox[0] - b);
    y0_ = (float)(w.bbox[1] - b);
    w_ = (int)std::ceil((W + 2 * b) / b) + 1;
    h_ = (int)std::ceil((H + 2 * b) / b) + 1;
    const int nb = w_ * h_, nc = 2 * nE;
    const double slack = 1e-3 * b + 1e-3;
    std::vector<int> ax(nc), bx(nc), ay(nc), by(nc);
    start_.assign(nb + 1, 0);
    for (int c = 0; c < nc; c++) {
        int e = c >> 1, side = c & 1;
        P2f q[4] = {Vert(w, w.ecm.edge_v[2 * e]), Cl(w, e, side), Cl(w, e, 2 + side), Vert(w, w.ecm.edge_v[2 * e + 1])};
        double lox = q[0].x, hix = q[0].x, loy = q[0].y, hiy = q[0].y;
        for (int k = 1; k < 4; k++) {
            lox = std::min<double>(lox, q[k].x); hix = std::max<double>(hix, q[k].x);
            loy = std::min<double>(loy, q[k].y); hiy = std::max<double>(hiy, q[k].y);
        }
        ax[c] = std::max(0, (int)std::floor((lox - slack - x0_) / b));
        bx[c] = std::min(w_ - 1, (int)std::floor((hix + slack - x0_) / b));
        ay[c] = std::max(0, (int)std::floor((loy - slack - y0_) / b));
        by[c] = std::min(h_ - 1, (int)std::floor((hiy + slack - y0_) / b));
        for (int y = ay[c]; y <= by[c]; y++)
            for (int x = ax[c]; x <= bx[c]; x++) start_[y * w_ + x + 1]++;
    }
    for (int i = 0; i < nb; i++) start_[i + 1] += start_[i];
    items_.resize(start_[nb]);
    std::vector<int> fill(start_.begin(), start_.end() - 1);
    for (int c = 0; c < nc; c++)
        for (int y = ay[c]; y <= by[c]; y++)
            for (int x = ax[c]; x <= bx[c]; x++) items_[fill[y * w_ + x]++] = c;
    // exact-level points: the y of every cell-polygon vertex, and per bin row the cells reaching into it
    levels_.clear();
    for (int c = 0; c < nc; c++) {
        const int e = c >> 1, side = c & 1;
        const float ys[4] = {Vert(w, w.ecm.edge_v[2 * e]).y, Cl(w, e, side).y, Cl(w, e, 2 + side).y, Vert(w, w.ecm.edge_v[2 * e + 1]).y};
        pass_through_levels(ys, levels_);
    }
    std::sort(levels_.begin(), levels_.end());
    levels_.erase(std::unique(levels_.begin(), levels_.end()), levels_.end());
    row_start_.assign(h_ + 1, 0);
    for (int c = 0; c < nc; c++)
        for (int y = ay[c]; y <= by[c]; y++) row_start_[y + 1]++;
    for (int i = 0; i < h_; i++) row_start_[i + 1] += row_start_[i];
    row_items_.resize(row_start_[h_]);
    std::vector<int> rfill(row_start_.begin(), row_start_.end() - 1);
    for (int c = 0; c < nc; c++)
        for (int y = ay[c]; y <= by[c]; y++) row_items_[rfill[y]++] = c;
}

int CellLocator::FindCell(const FlatWorld& w, float x, float y) const {
    const P2f p = P(x, y);
    const float fx = (x - x0_) / bin_, fy = (y - y0_) / bin_;
    if (std::binary_search(levels_.begin(), levels_.end(), y)) {  // exactly level with a cell vertex (see planner.h)
        if (fy >= 0.0f && fy < (float)h_) {
            const int r = (int)fy;
            for (int i = row_start_[r]; i < row_start_[r + 1]; i++)
                if (CellContains(w, row_items_[i], p)) return row_items_[i];
            return -1;
        }
    } else if (fx >= 0.0f && fy >= 0.0f && fx < (float)w_ && fy < (float)h_) {
        const int b = (int)fy * w_ + (int)fx;
        for (int i = start_[b]; i < start_[b + 1]; i++)
            if (CellContains(w, items_[i], p)) return items_[i];
        return -1;
    }
    for (int c = 0; c < 2 * w.ecm.num_edges(); c++)
        if (CellContains(w, c, p)) return c;
    return -1;
}

bool RetractPoint(const FlatWorld& w, const CellLocator& loc, P2f location, P2f& out, int& out_edge) {
    const int cell = loc.FindCell(w, location.x, location.y);
    if (cell < 0) return false;
    const int e = cell >> 1;
    out_edge = e;
    const P2f p1 = Vert(w, w.ecm.edge_v[2 * e]), p2 = Vert(w, w.ecm.edge_v[2 * e + 1]);
    const bool left = IsLeftOfSegment(p1, p2, location);
    const P2f o1 = Cl(w, e, left ? 0 : 1), o2 = Cl(w, e, left ? 2 : 3);
    P2f ray;
    if (Approx(o1, o2)) {
        ray = (p1 - o1) + (p2 - o1);
    } else {
        P2f v = o2 - o1;
        ray = left ? P(v.y, -v.x) : P(-v.y, v.x);
    }
    ray = Normalized(ray);
    // GetRayToLineSegmentIntersection (UtilityFunctions.cpp:323-349)
    const P2f v1 = location - p1, v2 = p2 - p1, v3 = P(-ray.y, ray.x);
    const float dot = v2.x * v3.x + v2.y * v3.y;
    if (std::fabs(dot) < 0.000001) return false;
    const float t1 = (v2.x * v1.y - v2.y * v1.x) / dot;
    const float t2 = (v1.x * v3.x + v1.y * v3.y) / dot;
    if (t1 >= 0.0 && (t2 >= 0.0 && t2 <= 1.0)) {
        out = P(location.x + ray.x * t1, location.y + ray.y * t1);
        return true;
    }
    return false;
}

// ------------------------------------------------------------------------------------------------
PathPlanner::PathPlanner(const FlatWorld* world) : w_(world) {
    owned_loc_ = new CellLocator();
    owned_loc_->Build(*world);
    loc_ = owned_loc_;
    const int nV = world->ecm.num_vertices();
    g_.assign(nV, kMaxFloat);
    f_.assign(nV, kMaxFloat);
    parent_.assign(nV, nV);
    visited_.assign(nV, 0);
}
PathPlanner::PathPlanner(const FlatWorld* world, const CellLocator* shared) : w_(world), loc_(shared) {
    const int nV = world->ecm.num_vertices();
    g_.assign(nV, kMaxFloat);
    f_.assign(nV, kMaxFloat);
    parent_.assign(nV, nV);
    visited_.assign(nV, 0);
}
PathPlanner::~PathPlanner() { delete owned_loc_; }

// AStar::FindPath (AStar.cpp:44-160).  The open list is a std::priority_queue of node references
// ordered by the node's CURRENT fCost (AStarCompare, AStar.h:30-37): costs are updated after the
// push and while older entries of the same node are still queued, exactly like the reference.
bool PathPlanner::AStar(P2f startLoc, P2f goalLoc, int startEdge, int goalEdge, float clearance, std::vector<int>& outPath) {
    const FlatWorld& w = *w_;
    const int INVALID = w.ecm.num_vertices();
    auto cmp = [this](int a, int b) { return f_[a] > f_[b]; };
    std::priority_queue<int, std::vector<int>, decltype(cmp)> open(cmp);
    auto touch = [this](int v) { touched_.push_back(v); };
    auto clean = [this, INVALID]() {  // CleanRequestData (AStar.cpp:162-176)
        for (int v : touched_) { f_[v] = kMaxFloat; g_[v] = kMaxFloat; parent_[v] = INVALID; visited_[v] = 0; }
        touched_.clear();
    };
    const int sa = w.ecm.edge_v[2 * startEdge + 1];  // half_edges[0].v_target_idx
    const int sb = w.ecm.edge_v[2 * startEdge];      // half_edges[1].v_target_idx
    touch(sa); touch(sb);
    g_[sa] = Distance(startLoc, Vert(w, sa));
    f_[sa] = g_[sa] + Distance(Vert(w, sa), goalLoc);
    g_[sb] = Distance(startLoc, Vert(w, sb));
    f_[sb] = g_[sb] + Distance(Vert(w, sb), goalLoc);
    open.push(sa);
    open.push(sb);
    const int ga = w.ecm.edge_v[2 * goalEdge + 1], gb = w.ecm.edge_v[2 * goalEdge];
    while (!open.empty()) {
        while (!open.empty() && visited_[open.top()]) open.pop();
        if (open.empty()) { clean(); return false; }
        const int cur = open.top();
        open.pop();
        visited_[cur] = 1;
        if (w.ecm.vert_clear[cur] < clearance) continue;
        if (cur == ga || cur == gb) {  // ConstructPath (AStar.cpp:184-212)
            std::vector<int> rev;
            if (cur == ga) { rev.push_back(gb); rev.push_back(ga); }
            else { rev.push_back(ga); rev.push_back(gb); }
            int nxt = parent_[cur];
            while (nxt < INVALID) { rev.push_back(nxt); nxt = parent_[nxt]; }
            for (int i = (int)rev.size() - 1; i >= 0; i--) outPath.push_back(rev[i]);
            clean();
            return true;
        }
        // neighbour ring (AStar.cpp:104-152): the first walk only locates the ring start again
        int he = w.ecm.vert_he[cur];
        const int startNb = HeTarget(w, he);
        int nb = startNb;
        do {
            he = w.ecm.he_next[he];
            nb = HeTarget(w, he);
        } while (startNb != nb);
        do {
            if (visited_[nb]) {
                he = w.ecm.he_next[he];
                nb = HeTarget(w, he);
                continue;
            }
            open.push(nb);
            touch(nb);
            const float newG = g_[cur] + Distance(Vert(w, cur), Vert(w, nb));
            if (newG < g_[nb]) {
                const float newF = newG + Distance(Vert(w, nb), goalLoc);
                parent_[nb] = cur;
                f_[nb] = newF;
                g_[nb] = newG;
            }
            he = w.ecm.he_next[he];
            nb = HeTarget(w, he);
        } while (nb != startNb);
    }
    clean();
    return false;
}

namespace {

// ECMPathPlanner::SampleCorridorArc (ECMPathPlanner.cpp:256-278)
void SampleCorridorArc(P2f p1, P2f p2, P2f o1, P2f o2, P2f c, float radius, bool leftArc, std::vector<Seg>& portals) {
    const float maxCurveSampleLength = 10.0f;
    if (leftArc) portals.push_back(Seg{p1, o1});
    else portals.push_back(Seg{o1, p1});
    const float edgeLength = Length(p2 - p1);
    const int numSamples = (int)std::ceil(edgeLength / maxCurveSampleLength);
    const float sampleLength = edgeLength / (float)numSamples;
    const P2f edgeDirection = (p2 - p1) / edgeLength;
    for (int i = 0; i < numSamples; i++) {
        P2f p = p1 + edgeDirection * sampleLength * (float)i;
        P2f arcDirection = Normalized(p - c);
        p = c + arcDirection * radius;
        if (leftArc) portals.push_back(Seg{p, o2});
        else portals.push_back(Seg{o2, p});
    }
}

// ECMPathPlanner::Funnel (ECMPathPlanner.cpp:323-411).  The reference loops forever (until
// bad_alloc) when Point::Approximate(p, p) is false, which happens for |coordinate| >= 2048 where
// half a float ulp exceeds EPSILON; we bound the number of scan steps and report failure instead.
bool Funnel(const std::vector<Seg>& portals, P2f start, P2f goal, std::vector<P2f>& out) {
    long budget = 64L * (long)portals.size() + 1024;
    P2f portalApex = start, portalLeft = start, portalRight = start;
    int leftIdx = 0, rightIdx = 0, apexIdx = 0;
    out.push_back(start);
    for (int i = 0; i < (int)portals.size(); i++) {
        if (--budget < 0) return false;
        const P2f left = portals[i].p0, right = portals[i].p1;
        if (TriangleArea(portalApex, portalRight, right) <= 0.0f) {
            if (Approx(portalApex, portalRight) || TriangleArea(portalApex, portalLeft, right) > 0.0f) {
                portalRight = right;
                rightIdx = i;
            } else {
                out.push_back(portalLeft);
                portalApex = portalLeft;
                apexIdx = leftIdx;
                portalLeft = portalApex;
                portalRight = portalApex;
                leftIdx = apexIdx;
                rightIdx = apexIdx;
                i = apexIdx;
                continue;
            }
        }
        if (TriangleArea(portalApex, portalLeft, left) >= 0.0f) {
            if (Approx(portalApex, portalLeft) || TriangleArea(portalApex, portalRight, left) < 0.0f) {
                portalLeft = left;
                leftIdx = i;
            } else {
                out.push_back(portalRight);
                portalApex = portalRight;
                apexIdx = rightIdx;
                portalLeft = portalApex;
                portalRight = portalApex;
                leftIdx = apexIdx;
                rightIdx = apexIdx;
                i = apexIdx;
                continue;
            }
        }
    }
    if (!Approx(out.back(), goal)) out.push_back(goal);
    return true;
}

}  // namespace

bool PathPlanner::FindPath(P2f start, P2f goal, float clearance, std::vector<P2f>& outPath) {
    const FlatWorld& w = *w_;
    outPath.clear();
    clearance += 0.0f;  // preferredAdditionalClearance (Simulator.cpp:108)
    // 1. cells (ECMPathPlanner.cpp:45-52)
    if (loc_->FindCell(w, start.x, start.y) < 0 || loc_->FindCell(w, goal.x, goal.y) < 0) return false;
    // 2. retraction (ECMPathPlanner.cpp:58-70)
    P2f retrStart, retrGoal;
    int startEdge = -1, goalEdge = -1;
    if (!RetractPoint(w, *loc_, start, retrStart, startEdge)) return false;
    if (!RetractPoint(w, *loc_, goal, retrGoal, goalEdge)) return false;
    if (startEdge == goalEdge) {  // ECMPathPlanner.cpp:74-80
        outPath.push_back(start);
        outPath.push_back(goal);
        return true;
    }
    // 3. A* on the medial axis (ECMPathPlanner.cpp:84-89)
    std::vector<int> astar;
    if (!AStar(retrStart, retrGoal, startEdge, goalEdge, clearance, astar)) return false;
    // half-edge path (ECMPathPlanner.cpp:93-113)
    std::vector<int> edgePath;
    for (int i = 0; i + 1 < (int)astar.size(); i++) {
        const int i1 = astar[i], i2 = astar[i + 1];
        int he = w.ecm.vert_he[i1];
        const int heStart = he;
        do {
            if (HeTarget(w, he) == i2) { edgePath.push_back(he); break; }
            he = w.ecm.he_next[he];
        } while (heStart != he);
    }
    if (edgePath.empty()) return false;  // the reference would index an empty corridor (UB); cannot happen on connected graphs
    // 4. corridor (ECMPathPlanner.cpp:146-158) and its shrunk bounds (:160-214)
    const int m = (int)edgePath.size();
    std::vector<P2f> centers(m), lb(m), rb(m), lcb, rcb;
    std::vector<float> radii(m);
    std::vector<int> curve;  // 0 LINEAR, 1 LEFT_ARC, 2 RIGHT_ARC
    for (int i = 0; i < m; i++) {
        const int src = HeSource(w, edgePath[i]);
        centers[i] = Vert(w, src);
        radii[i] = w.ecm.vert_clear[src];
        lb[i] = HeClosestLeft(w, edgePath[i]);
        rb[i] = HeClosestRight(w, edgePath[i]);
    }
    auto shrink = [&](int i) {
        if (radii[i] < clearance) {
            lcb.push_back(centers[i]);
            rcb.push_back(centers[i]);
        } else {
            P2f ml = Normalized(centers[i] - lb[i]);
            P2f mr = Normalized(centers[i] - rb[i]);
            lcb.push_back(lb[i] + ml * clearance);
            rcb.push_back(rb[i] + mr * clearance);
        }
    };
    for (int i = 0; i < m - 1; i++) {
        shrink(i);
        if (Approx(lb[i], lb[i + 1])) curve.push_back(1);
        else if (Approx(rb[i], rb[i + 1])) curve.push_back(2);
        else curve.push_back(0);
    }
    shrink(m - 1);
    // 5. portals (ECMPathPlanner.cpp:216-254)
    std::vector<Seg> portals;
    for (int i = 0; i < m - 1; i++) {
        switch (curve[i]) {
            case 0:
                portals.push_back(Seg{lcb[i], rcb[i]});
                portals.push_back(Seg{lcb[i + 1], rcb[i]});
                break;
            case 1:
                SampleCorridorArc(lcb[i], lcb[i + 1], rcb[i], rcb[i + 1], lb[i], clearance, true, portals);
                break;
            case 2:
                SampleCorridorArc(rcb[i], rcb[i + 1], lcb[i], lcb[i + 1], rb[i], clearance, false, portals);
                break;
        }
    }
    portals.push_back(Seg{lcb.back(), rcb.back()});
    {  // FitPortalRange (ECMPathPlanner.cpp:280-320)
        int first = 0;
        for (int i = 0; i < (int)portals.size(); i++)
            if (!IsLeftOfSegment(portals[i].p0, portals[i].p1, start)) { first = i; break; }
        portals.erase(portals.begin(), portals.begin() + first);
        int last = (int)portals.size() - 1;
        for (int i = (int)portals.size() - 1; i >= 0; i--)
            if (IsLeftOfSegment(portals[i].p0, portals[i].p1, goal)) { last = i; break; }
        const int toRemove = (int)portals.size() - last;
        for (int i = 0; i < toRemove && !portals.empty(); i++) portals.pop_back();
    }
    portals.push_back(Seg{goal, goal});
    // 7. funnel (ECMPathPlanner.cpp:127-133)
    if (!Funnel(portals, start, goal, outPath)) {
        outPath.clear();
        return false;
    }
    return true;
}

int PlanPaths(const FlatWorld& w, int n, const float* start_xy, const float* goal_xy, const float* clearance, int threads,
              std::vector<int>& out_off, std::vector<float>& out_xy) {
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    threads = std::max(1, std::min(threads, std::max(1, n / 64)));
    CellLocator loc;
    loc.Build(w);
    std::vector<std::vector<float>> part(threads);
    std::vector<std::vector<int>> lens(threads);
    std::vector<int> ok(threads, 0);
    auto work = [&](int t) {
        PathPlanner pl(&w, &loc);
        const int lo = (int)((long long)n * t / threads), hi = (int)((long long)n * (t + 1) / threads);
        std::vector<P2f> path;
        for (int i = lo; i < hi; i++) {
            bool good = pl.FindPath(P2f{start_xy[2 * i], start_xy[2 * i + 1]}, P2f{goal_xy[2 * i], goal_xy[2 * i + 1]}, clearance[i], path);
            if (!good) path.clear();
            ok[t] += good ? 1 : 0;
            lens[t].push_back((int)path.size());
            for (const P2f& p : path) { part[t].push_back(p.x); part[t].push_back(p.y); }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    out_off.assign(1, 0);
    out_xy.clear();
    int total_ok = 0;
    for (int t = 0; t < threads; t++) {
        for (int l : lens[t]) out_off.push_back(out_off.back() + l);
        out_xy.insert(out_xy.end(), part[t].begin(), part[t].end());
        total_ok += ok[t];
    }
    return total_ok;
}

}  // namespace ecmb200
