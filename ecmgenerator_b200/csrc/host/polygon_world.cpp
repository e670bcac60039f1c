#include "polygon_world.h"

#include <algorithm>
#include <cmath>
#include <map>

namespace ecmb200 {
namespace {

struct P2 {
    double x, y;
};
inline P2 operator+(P2 a, P2 b) { return {a.x + b.x, a.y + b.y}; }
inline P2 operator-(P2 a, P2 b) { return {a.x - b.x, a.y - b.y}; }
inline P2 operator*(P2 a, double s) { return {a.x * s, a.y * s}; }
inline double dot(P2 a, P2 b) { return a.x * b.x + a.y * b.y; }
inline double cross(P2 a, P2 b) { return a.x * b.y - a.y * b.x; }
inline double norm(P2 a) { return std::sqrt(dot(a, a)); }

struct Site {
    bool is_point = false;
    P2 a{0, 0}, b{0, 0};  // a point site uses `a`
    P2 dir{0, 0}, nrm{0, 0};
    double len = 0, c = 0;  // segment: unit direction, left unit normal, nrm . x = c on its line
    int end0 = -1, end1 = -1;  // segment: the point sites at its ends
};

struct Gen {
    std::vector<Site> sites;
    std::vector<std::vector<P2>> polys;
    double x0, y0, x1, y1, scale, tol;

    double dist(const Site& s, P2 x) const {
        if (s.is_point) return norm(x - s.a);
        const double t = dot(x - s.a, s.dir);
        if (t <= 0) return norm(x - s.a);
        if (t >= s.len) return norm(x - s.b);
        return std::fabs(dot(s.nrm, x) - s.c);
    }
    P2 closest(const Site& s, P2 x) const {
        if (s.is_point) return s.a;
        const double t = std::min(std::max(dot(x - s.a, s.dir), 0.0), s.len);
        return s.a + s.dir * t;
    }
    // foot of x on a segment site strictly inside it (with slack `e` at both ends)
    bool foot_inside(const Site& s, P2 x, double e) const {
        if (s.is_point) return true;
        const double t = dot(x - s.a, s.dir);
        return t > e && t < s.len - e;
    }
    bool in_free_space(P2 x) const {
        if (x.x < x0 - tol || x.x > x1 + tol || x.y < y0 - tol || x.y > y1 + tol) return false;
        for (const auto& poly : polys) {
            bool inside = false;
            double dmin = 1e300;
            const size_t m = poly.size();
            for (size_t i = 0, j = m - 1; i < m; j = i++) {
                const P2 p = poly[i], q = poly[j];
                if ((p.y > x.y) != (q.y > x.y) && x.x < (q.x - p.x) * (x.y - p.y) / (q.y - p.y) + p.x) inside = !inside;
                const P2 d = q - p;
                const double l2 = dot(d, d);
                const double t = l2 > 0 ? std::min(1.0, std::max(0.0, dot(x - p, d) / l2)) : 0.0;
                dmin = std::min(dmin, norm(x - (p + d * t)));
            }
            if (inside && dmin > 10 * tol) return false;  // strictly inside an obstacle; its boundary is free space's rim
        }
        return true;
    }
    // nearest-site test: nobody closer than r - slack, apart from the listed sites
    bool empty_circle(P2 x, double r, double slack) const {
        for (const Site& s : sites)
            if (dist(s, x) < r - slack) return false;
        return true;
    }
};

struct Vertex {
    P2 p;
    double r;
    std::vector<int> sites;
};

// All points equidistant from the three sites (see polygon_world.h): a line in (x, y, r) from the linear equations, cut
// with the quadric of a point site where there is one.
void solve_triple(const Gen& g, int i, int j, int k, std::vector<std::pair<P2, double>>& out) {
    const int idx[3] = {i, j, k};
    std::vector<int> segs, pts;
    for (int t : idx) (g.sites[t].is_point ? pts : segs).push_back(t);
    const int ns = (int)segs.size();
    for (int mask = 0; mask < (1 << ns); mask++) {
        double A[3][3], B[3];
        int rows = 0;
        for (int q = 0; q < ns; q++) {
            const Site& s = g.sites[segs[q]];
            const double sg = (mask >> q & 1) ? -1.0 : 1.0;
            A[rows][0] = sg * s.nrm.x; A[rows][1] = sg * s.nrm.y; A[rows][2] = -1.0; B[rows] = sg * s.c;
            rows++;
        }
        for (size_t q = 1; q < pts.size(); q++) {
            const P2 p = g.sites[pts[0]].a, w = g.sites[pts[q]].a;
            A[rows][0] = 2 * (w.x - p.x); A[rows][1] = 2 * (w.y - p.y); A[rows][2] = 0.0; B[rows] = dot(w, w) - dot(p, p);
            rows++;
        }
        if (rows == 3) {  // SSS
            const double det = A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
                               A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
            if (std::fabs(det) < 1e-12) continue;
            auto det3 = [&](int col) {
                double M[3][3];
                for (int r = 0; r < 3; r++)
                    for (int c = 0; c < 3; c++) M[r][c] = c == col ? B[r] : A[r][c];
                return M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
                       M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
            };
            out.push_back({P2{det3(0) / det, det3(1) / det}, det3(2) / det});
            continue;
        }
        // two linear equations: X(t) = u + t w
        const double w[3] = {A[0][1] * A[1][2] - A[0][2] * A[1][1], A[0][2] * A[1][0] - A[0][0] * A[1][2], A[0][0] * A[1][1] - A[0][1] * A[1][0]};
        const double wn = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
        const double an = std::sqrt(A[0][0] * A[0][0] + A[0][1] * A[0][1] + A[0][2] * A[0][2]) * std::sqrt(A[1][0] * A[1][0] + A[1][1] * A[1][1] + A[1][2] * A[1][2]);
        if (wn < 1e-12 * an) continue;  // dependent equations
        // minimum-norm particular solution u = A^T (A A^T)^-1 B
        const double g00 = A[0][0] * A[0][0] + A[0][1] * A[0][1] + A[0][2] * A[0][2], g11 = A[1][0] * A[1][0] + A[1][1] * A[1][1] + A[1][2] * A[1][2];
        const double g01 = A[0][0] * A[1][0] + A[0][1] * A[1][1] + A[0][2] * A[1][2];
        const double gd = g00 * g11 - g01 * g01;
        const double l0 = (g11 * B[0] - g01 * B[1]) / gd, l1 = (g00 * B[1] - g01 * B[0]) / gd;
        const double u[3] = {A[0][0] * l0 + A[1][0] * l1, A[0][1] * l0 + A[1][1] * l1, A[0][2] * l0 + A[1][2] * l1};
        const double ww[3] = {w[0] / wn, w[1] / wn, w[2] / wn};
        const P2 p = g.sites[pts[0]].a;
        const double qa = ww[0] * ww[0] + ww[1] * ww[1] - ww[2] * ww[2];
        const double qb = 2 * ((u[0] - p.x) * ww[0] + (u[1] - p.y) * ww[1] - u[2] * ww[2]);
        const double qc = (u[0] - p.x) * (u[0] - p.x) + (u[1] - p.y) * (u[1] - p.y) - u[2] * u[2];
        double ts[2];
        int nt = 0;
        if (std::fabs(qa) < 1e-14) {
            if (std::fabs(qb) > 1e-14) ts[nt++] = -qc / qb;
        } else {
            double disc = qb * qb - 4 * qa * qc;
            if (disc < 0 && disc > -1e-9 * (qb * qb + std::fabs(4 * qa * qc))) disc = 0;
            if (disc >= 0) {
                const double sq = std::sqrt(disc);
                const double q = -0.5 * (qb + (qb >= 0 ? sq : -sq));
                ts[nt++] = q / qa;
                if (q != 0) ts[nt++] = qc / q; else ts[nt++] = 0;
            }
        }
        for (int t = 0; t < nt; t++) out.push_back({P2{u[0] + ts[t] * ww[0], u[1] + ts[t] * ww[1]}, u[2] + ts[t] * ww[2]});
    }
}

}  // namespace

bool BuildPolygonWorld(const float bbox[4], int n_polys, const int* poly_first, const float* poly_xy, FlatWorld& out, std::string* error) {
    auto fail = [&](const char* m) { if (error) *error = m; return false; };
    Gen g;
    g.x0 = bbox[0]; g.y0 = bbox[1]; g.x1 = bbox[2]; g.y1 = bbox[3];
    if (!(g.x1 > g.x0) || !(g.y1 > g.y0)) return fail("empty walkable area");
    g.scale = std::hypot(g.x1 - g.x0, g.y1 - g.y0);
    g.tol = 1e-9 * g.scale;
    if (n_polys < 0 || (n_polys > 0 && (!poly_first || !poly_xy))) return fail("bad polygon arrays");
    // ---- sites: the end points first (shared ones once), then the segments in union order (Environment.cpp:188-222)
    std::vector<std::pair<P2, P2>> segs;
    segs.push_back({{g.x0, g.y0}, {g.x1, g.y0}}); segs.push_back({{g.x1, g.y0}, {g.x1, g.y1}});
    segs.push_back({{g.x0, g.y1}, {g.x1, g.y1}}); segs.push_back({{g.x0, g.y1}, {g.x0, g.y0}});
    for (int k = 0; k < n_polys; k++) {
        const int a = poly_first[k], b = poly_first[k + 1];
        if (b - a < 3) return fail("an obstacle needs at least 3 vertices");
        std::vector<P2> poly;
        double area2 = 0;
        for (int i = a; i < b; i++) {
            poly.push_back({poly_xy[2 * i], poly_xy[2 * i + 1]});
            const int j = i + 1 < b ? i + 1 : a;
            area2 += (double)poly_xy[2 * i] * poly_xy[2 * j + 1] - (double)poly_xy[2 * j] * poly_xy[2 * i + 1];
            const P2 p = poly.back();
            if (!(p.x > g.x0 && p.x < g.x1 && p.y > g.y0 && p.y < g.y1)) return fail("obstacles must lie strictly inside the walkable area");
        }
        if (!(area2 > 0)) return fail("obstacle vertices must be counter-clockwise (Environment.cpp:198)");
        for (size_t i = 0; i < poly.size(); i++) segs.push_back({poly[i], poly[(i + 1) % poly.size()]});
        g.polys.push_back(poly);
    }
    std::vector<P2> pts;
    auto point_id = [&](P2 p) {
        for (size_t i = 0; i < pts.size(); i++)
            if (norm(pts[i] - p) <= g.tol) return (int)i;
        pts.push_back(p);
        return (int)pts.size() - 1;
    };
    std::vector<std::pair<int, int>> seg_ends;
    for (auto& s : segs) {
        if (norm(s.second - s.first) <= 1e-7 * g.scale) return fail("zero-length obstacle edge");
        seg_ends.push_back({point_id(s.first), point_id(s.second)});
    }
    const int np = (int)pts.size();
    for (const P2& p : pts) { Site s; s.is_point = true; s.a = s.b = p; g.sites.push_back(s); }
    for (size_t i = 0; i < segs.size(); i++) {
        Site s;
        s.a = segs[i].first; s.b = segs[i].second;
        s.len = norm(s.b - s.a);
        s.dir = (s.b - s.a) * (1.0 / s.len);
        s.nrm = {-s.dir.y, s.dir.x};
        s.c = dot(s.nrm, s.a);
        s.end0 = seg_ends[i].first; s.end1 = seg_ends[i].second;
        g.sites.push_back(s);
    }
    const int n = (int)g.sites.size();
    if (n > 1200) return fail("too many sites for the direct construction (use the lattice generator for city maps)");
    // ---- Voronoi vertices
    std::vector<Vertex> verts;
    const double merge = 1e-7 * g.scale, tie = 1e-7 * g.scale;
    std::vector<std::pair<P2, double>> cand;
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++)
            for (int k = j + 1; k < n; k++) {
                cand.clear();
                solve_triple(g, i, j, k, cand);
                for (auto& c : cand) {
                    const P2 x = c.first;
                    double r = c.second;
                    if (!(r > -tie) || !std::isfinite(r) || !std::isfinite(x.x) || !std::isfinite(x.y)) continue;
                    r = std::max(r, 0.0);
                    if (std::fabs(g.dist(g.sites[i], x) - r) > tie || std::fabs(g.dist(g.sites[j], x) - r) > tie || std::fabs(g.dist(g.sites[k], x) - r) > tie) continue;
                    if (!g.empty_circle(x, r, tie) || !g.in_free_space(x)) continue;
                    bool dup = false;
                    for (const Vertex& v : verts)
                        if (norm(v.p - x) <= merge) { dup = true; break; }
                    if (dup) continue;
                    Vertex v;
                    v.p = x; v.r = r;
                    for (int s = 0; s < n; s++)
                        if (g.dist(g.sites[s], x) <= r + tie) v.sites.push_back(s);
                    verts.push_back(v);
                }
            }
    // ---- edges: consecutive vertices along the bisector of every pair of sites
    struct Edge { int v0, v1, left, right; };
    std::vector<Edge> edges;
    std::map<std::pair<int, int>, std::vector<int>> by_pair;
    for (int v = 0; v < (int)verts.size(); v++)
        for (size_t a = 0; a < verts[v].sites.size(); a++)
            for (size_t b = a + 1; b < verts[v].sites.size(); b++) by_pair[{verts[v].sites[a], verts[v].sites[b]}].push_back(v);
    auto edge_between = [&](int sa, int sb, P2 m) {  // is the bisector point m a point of the Voronoi edge (sa, sb)?
        const double ra = g.dist(g.sites[sa], m), rb = g.dist(g.sites[sb], m);
        if (std::fabs(ra - rb) > 1e-6 * g.scale) return false;
        const double end_slack = 1e-7 * g.scale;
        if (!g.foot_inside(g.sites[sa], m, end_slack) || !g.foot_inside(g.sites[sb], m, end_slack)) return false;
        return g.empty_circle(m, std::min(ra, rb), tie) && g.in_free_space(m);
    };
    for (auto& kv : by_pair) {
        const int sa = kv.first.first, sb = kv.first.second;
        std::vector<int>& vs = kv.second;
        if (vs.size() < 2) continue;
        const Site& A = g.sites[sa];
        const Site& B = g.sites[sb];
        if (A.is_point != B.is_point) {  // a segment with its own end point: Boost's secondary edge, not part of the ECM
            const Site& S = A.is_point ? B : A;
            const int pid = A.is_point ? sa : sb;
            if (S.end0 == pid || S.end1 == pid) continue;
        }
        auto try_chain = [&](std::vector<int> chain, P2 d, bool parabola, const Site* seg, P2 focus) {
            std::sort(chain.begin(), chain.end(), [&](int u, int v) { return dot(verts[u].p, d) < dot(verts[v].p, d); });
            for (size_t q = 0; q + 1 < chain.size(); q++) {
                const P2 u = verts[chain[q]].p, v = verts[chain[q + 1]].p;
                P2 m = (u + v) * 0.5;
                if (parabola) {  // the parabola point whose foot on the directrix lies midway between the two feet
                    const double su = dot(u - seg->a, seg->dir), sv = dot(v - seg->a, seg->dir), sm = 0.5 * (su + sv);
                    const double h = dot(seg->nrm, focus) - seg->c;  // signed height of the focus over the directrix
                    const double sf = dot(focus - seg->a, seg->dir);
                    if (std::fabs(h) < 1e-12 * g.scale) continue;
                    const double yy = ((sm - sf) * (sm - sf) + h * h) / (2 * h);
                    m = seg->a + seg->dir * sm + seg->nrm * yy;
                }
                if (!edge_between(sa, sb, m)) continue;
                Edge e;
                e.v0 = chain[q]; e.v1 = chain[q + 1];
                const P2 dd = v - u;
                const bool a_left = cross(dd, g.closest(A, m) - m) > 0;
                e.left = a_left ? sa : sb;
                e.right = a_left ? sb : sa;
                edges.push_back(e);
            }
        };
        if (A.is_point && B.is_point) {
            const P2 pq = B.a - A.a;
            try_chain(vs, P2{-pq.y, pq.x}, false, nullptr, P2{0, 0});
        } else if (A.is_point != B.is_point) {
            const Site& S = A.is_point ? B : A;
            const Site& F = A.is_point ? A : B;
            try_chain(vs, S.dir, true, &S, F.a);
        } else {
            // two segments: the vertices lie on (at most) the two angle bisectors of their lines, or on the mid line of parallel ones
            std::vector<P2> dirs;
            const P2 d1 = A.dir + B.dir, d2 = A.dir - B.dir;
            if (norm(d1) > 1e-9) dirs.push_back(d1 * (1.0 / norm(d1)));
            if (norm(d2) > 1e-9) dirs.push_back(d2 * (1.0 / norm(d2)));
            for (const P2& d : dirs) {
                // group the vertices by the line (direction d) they lie on
                const P2 nd{-d.y, d.x};
                std::vector<int> rest(vs);
                while (!rest.empty()) {
                    const double off = dot(verts[rest[0]].p, nd);
                    std::vector<int> line, other;
                    for (int v : rest) (std::fabs(dot(verts[v].p, nd) - off) <= 1e-6 * g.scale ? line : other).push_back(v);
                    if (line.size() >= 2) try_chain(line, d, false, nullptr, P2{0, 0});
                    rest.swap(other);
                }
            }
        }
    }
    // the same pair of vertices can come out twice (the two bisector directions of parallel segments coincide): once is enough
    std::sort(edges.begin(), edges.end(), [](const Edge& a, const Edge& b) { return std::make_pair(std::min(a.v0, a.v1), std::max(a.v0, a.v1)) < std::make_pair(std::min(b.v0, b.v1), std::max(b.v0, b.v1)); });
    edges.erase(std::unique(edges.begin(), edges.end(), [](const Edge& a, const Edge& b) { return std::min(a.v0, a.v1) == std::min(b.v0, b.v1) && std::max(a.v0, a.v1) == std::max(b.v0, b.v1); }), edges.end());
    if (edges.empty()) return fail("no medial axis found");
    // ---- flat world: vertices that carry an edge, ordered by (y, x); edges by their end points
    std::vector<int> used(verts.size(), 0);
    for (const Edge& e : edges) used[e.v0] = used[e.v1] = 1;
    std::vector<int> order;
    for (int v = 0; v < (int)verts.size(); v++)
        if (used[v]) order.push_back(v);
    std::sort(order.begin(), order.end(), [&](int a, int b) { return verts[a].p.y != verts[b].p.y ? verts[a].p.y < verts[b].p.y : verts[a].p.x < verts[b].p.x; });
    std::vector<int> new_id(verts.size(), -1);
    FlatWorld w;
    for (int i = 0; i < 4; i++) w.bbox[i] = bbox[i];
    for (size_t i = 0; i < order.size(); i++) {
        new_id[order[i]] = (int)i;
        w.ecm.vert_xy.push_back((float)verts[order[i]].p.x);
        w.ecm.vert_xy.push_back((float)verts[order[i]].p.y);
        w.ecm.vert_clear.push_back((float)verts[order[i]].r);
        w.ecm.vert_he.push_back(-1);
    }
    for (Edge& e : edges) {
        e.v0 = new_id[e.v0]; e.v1 = new_id[e.v1];
        if (e.v0 > e.v1) { std::swap(e.v0, e.v1); std::swap(e.left, e.right); }
    }
    std::sort(edges.begin(), edges.end(), [](const Edge& a, const Edge& b) { return a.v0 != b.v0 ? a.v0 < b.v0 : a.v1 < b.v1; });
    std::vector<std::vector<int>> out_he(order.size());
    for (size_t e = 0; e < edges.size(); e++) {
        const Edge& ed = edges[e];
        const P2 p0 = verts[order[ed.v0]].p, p1 = verts[order[ed.v1]].p;
        w.ecm.edge_v.push_back(ed.v0);
        w.ecm.edge_v.push_back(ed.v1);
        const P2 cl[4] = {g.closest(g.sites[ed.left], p0), g.closest(g.sites[ed.right], p0), g.closest(g.sites[ed.left], p1), g.closest(g.sites[ed.right], p1)};
        for (const P2& c : cl) { w.ecm.edge_cl.push_back((float)c.x); w.ecm.edge_cl.push_back((float)c.y); }
        w.ecm.he_next.push_back(-1);
        w.ecm.he_next.push_back(-1);
        out_he[ed.v0].push_back(2 * (int)e);
        out_he[ed.v1].push_back(2 * (int)e + 1);
    }
    for (size_t v = 0; v < order.size(); v++) {  // rings of outgoing half-edges, by angle (AStar.cpp:108-152 walks them)
        auto& hs = out_he[v];
        const P2 p = verts[order[v]].p;
        auto angle = [&](int he) {
            const int e = he >> 1, tgt = (he & 1) ? w.ecm.edge_v[2 * e] : w.ecm.edge_v[2 * e + 1];
            const P2 q = verts[order[tgt]].p;
            return std::atan2(q.y - p.y, q.x - p.x);
        };
        std::sort(hs.begin(), hs.end(), [&](int a, int b) { return angle(a) < angle(b); });
        w.ecm.vert_he[v] = hs[0];
        for (size_t k = 0; k < hs.size(); k++) w.ecm.he_next[hs[k]] = hs[(k + 1) % hs.size()];
    }
    for (int k = 0; k < n_polys; k++) AppendObstacle(w.obst, poly_xy + 2 * poly_first[k], poly_first[k + 1] - poly_first[k]);
    (void)np;
    out = std::move(w);
    return true;
}

}  // namespace ecmb200
