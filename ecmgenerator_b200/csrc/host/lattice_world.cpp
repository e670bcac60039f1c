#include "lattice_world.h"

#include <algorithm>
#include <cmath>

namespace ecmb200 {

void AppendObstacle(FlatObstacles& o, const float* xy, int n) {
    if (o.first.empty()) o.first.push_back(0);
    const int base = o.num_vertices();
    for (int i = 0; i < n; i++) {
        o.xy.push_back(xy[2 * i]);
        o.xy.push_back(xy[2 * i + 1]);
        // same wrap-around as Obstacle::Initialize (ECMDataTypes.cpp:36-45)
        o.prev.push_back(base + (i == 0 ? n - 1 : i - 1));
        o.next.push_back(base + (i == 0 ? 1 % n : (i + 1) % n));
        o.convex.push_back(1);
    }
    if (n > 2) {
        for (int i = 0; i < n; i++) {
            const int p = o.prev[base + i], q = o.next[base + i], c = base + i;
            // isConvex = det(prev - next, p - prev) >= 0   (ECMDataTypes.cpp:52-59), float arithmetic
            const float ax = o.xy[2 * p] - o.xy[2 * q], ay = o.xy[2 * p + 1] - o.xy[2 * q + 1];
            const float bx = o.xy[2 * c] - o.xy[2 * p], by = o.xy[2 * c + 1] - o.xy[2 * p + 1];
            const float m1 = ax * by, m2 = ay * bx;
            o.convex[c] = (m1 - m2) >= 0.0f ? 1 : 0;
        }
    }
    o.first.push_back(base + n);
}

namespace {

struct P2 {
    double x, y;
};

struct Builder {
    FlatECM& g;
    std::vector<std::vector<int>> out_he;  // outgoing half-edges per vertex

    explicit Builder(FlatECM& ecm) : g(ecm) {}

    int vertex(P2 p, double clearance) {
        g.vert_xy.push_back((float)p.x);
        g.vert_xy.push_back((float)p.y);
        g.vert_clear.push_back((float)clearance);
        g.vert_he.push_back(-1);
        out_he.emplace_back();
        return g.num_vertices() - 1;
    }
    P2 pos(int v) const { return {g.vert_xy[2 * v], g.vert_xy[2 * v + 1]}; }

    int edge(int v0, int v1, P2 L0, P2 R0, P2 L1, P2 R1) {
        const int e = g.num_edges();
        g.edge_v.push_back(v0);
        g.edge_v.push_back(v1);
        const P2 pts[4] = {L0, R0, L1, R1};
        for (const P2& p : pts) {
            g.edge_cl.push_back((float)p.x);
            g.edge_cl.push_back((float)p.y);
        }
        g.he_next.push_back(-1);
        g.he_next.push_back(-1);
        out_he[v0].push_back(2 * e);      // half-edge 0 leaves v0
        out_he[v1].push_back(2 * e + 1);  // half-edge 1 leaves v1
        return e;
    }

    // Straight street piece v0 -> v1 between two parallel walls at distance h on either side.
    int street(int v0, int v1, double h) {
        const P2 a = pos(v0), b = pos(v1);
        const double dx = b.x - a.x, dy = b.y - a.y, l = std::sqrt(dx * dx + dy * dy);
        const double nx = -dy / l, ny = dx / l;  // left normal
        return edge(v0, v1, {a.x + nx * h, a.y + ny * h}, {a.x - nx * h, a.y - ny * h},
                    {b.x + nx * h, b.y + ny * h}, {b.x - nx * h, b.y - ny * h});
    }
    // Point/point bisector v0 -> v1 between the point sites `l` (left) and `r` (right).
    int point_bisector(int v0, int v1, P2 l, P2 r) { return edge(v0, v1, l, r, l, r); }
    // Dead-end diagonal from T into wall corner K; fa / fb are the feet of T on the two walls.
    int diagonal(int t, int k, P2 fa, P2 fb) {
        const P2 T = pos(t), K = pos(k);
        const double dx = K.x - T.x, dy = K.y - T.y;
        const double side_a = (-dy) * (fa.x - T.x) + dx * (fa.y - T.y);
        const P2 l = side_a > 0 ? fa : fb, r = side_a > 0 ? fb : fa;
        return edge(t, k, l, r, K, K);
    }

    void link_rings() {
        for (int v = 0; v < g.num_vertices(); v++) {
            auto& hs = out_he[v];
            const P2 p = pos(v);
            auto angle = [&](int he) {
                const int e = he >> 1, tgt = (he & 1) ? g.edge_v[2 * e] : g.edge_v[2 * e + 1];
                const P2 q = pos(tgt);
                return std::atan2(q.y - p.y, q.x - p.x);
            };
            std::sort(hs.begin(), hs.end(), [&](int a, int b) { return angle(a) < angle(b); });
            g.vert_he[v] = hs[0];
            for (size_t k = 0; k < hs.size(); k++) g.he_next[hs[k]] = hs[(k + 1) % hs.size()];
        }
    }
};

}  // namespace

bool BuildLatticeWorld(int nbx, const float* bx, int nby, const float* by, float W, float x0, float y0,
                       FlatWorld& out) {
    if (nbx < 1 || nby < 1 || !(W > 0.0f)) return false;
    if (nbx == 1 && nby == 1) return false;  // no street at all
    const double h = 0.5 * (double)W;
    for (int i = 0; i < nbx; i++)
        if (!(bx[i] >= h)) return false;
    for (int j = 0; j < nby; j++)
        if (!(by[j] >= h)) return false;

    // block lower/upper coordinates and street centres
    std::vector<double> bxl(nbx), bxh(nbx), byl(nby), byh(nby);
    double c = x0;
    for (int i = 0; i < nbx; i++) { bxl[i] = c; c += bx[i]; bxh[i] = c; c += W; }
    const double x1 = bxh[nbx - 1];
    c = y0;
    for (int j = 0; j < nby; j++) { byl[j] = c; c += by[j]; byh[j] = c; c += W; }
    const double y1 = byh[nby - 1];
    const int nsx = nbx - 1, nsy = nby - 1;  // vertical / horizontal street counts
    std::vector<double> cx(nsx), cy(nsy);
    for (int i = 0; i < nsx; i++) cx[i] = bxh[i] + h;
    for (int j = 0; j < nsy; j++) cy[j] = byh[j] + h;

    FlatWorld w;
    w.bbox[0] = x0; w.bbox[1] = y0; w.bbox[2] = (float)x1; w.bbox[3] = (float)y1;
    Builder b(w.ecm);
    const double cc = (double)W / std::sqrt(2.0);

    // crossings: centre + E, N, W, S mouths
    std::vector<int> C(nsx * nsy), ME(nsx * nsy), MN(nsx * nsy), MW(nsx * nsy), MS(nsx * nsy);
    for (int j = 0; j < nsy; j++)
        for (int i = 0; i < nsx; i++) {
            const int k = j * nsx + i;
            const double x = cx[i], y = cy[j];
            C[k] = b.vertex({x, y}, cc);
            ME[k] = b.vertex({x + h, y}, h);
            MN[k] = b.vertex({x, y + h}, h);
            MW[k] = b.vertex({x - h, y}, h);
            MS[k] = b.vertex({x, y - h}, h);
            const P2 ne{x + h, y + h}, nw{x - h, y + h}, sw{x - h, y - h}, se{x + h, y - h};
            b.point_bisector(C[k], ME[k], ne, se);
            b.point_bisector(C[k], MN[k], nw, ne);
            b.point_bisector(C[k], MW[k], sw, nw);
            b.point_bisector(C[k], MS[k], se, sw);
        }

    // horizontal streets, travelling +x
    for (int j = 0; j < nsy; j++) {
        const double y = cy[j];
        const int tw = b.vertex({x0 + h, y}, h), te = b.vertex({x1 - h, y}, h);
        const int aw = b.vertex({x0, y + h}, 0), bw = b.vertex({x0, y - h}, 0);
        const int ae = b.vertex({x1, y + h}, 0), be = b.vertex({x1, y - h}, 0);
        int prev = tw;
        for (int i = 0; i < nsx; i++) {
            b.street(prev, MW[j * nsx + i], h);
            prev = ME[j * nsx + i];
        }
        b.street(prev, te, h);
        b.diagonal(tw, aw, {x0, y}, {x0 + h, y + h});
        b.diagonal(tw, bw, {x0, y}, {x0 + h, y - h});
        b.diagonal(te, ae, {x1, y}, {x1 - h, y + h});
        b.diagonal(te, be, {x1, y}, {x1 - h, y - h});
    }
    // vertical streets, travelling +y
    for (int i = 0; i < nsx; i++) {
        const double x = cx[i];
        const int ts = b.vertex({x, y0 + h}, h), tn = b.vertex({x, y1 - h}, h);
        const int as = b.vertex({x - h, y0}, 0), bs = b.vertex({x + h, y0}, 0);
        const int an = b.vertex({x - h, y1}, 0), bn = b.vertex({x + h, y1}, 0);
        int prev = ts;
        for (int j = 0; j < nsy; j++) {
            b.street(prev, MS[j * nsx + i], h);
            prev = MN[j * nsx + i];
        }
        b.street(prev, tn, h);
        b.diagonal(ts, as, {x, y0}, {x - h, y0 + h});
        b.diagonal(ts, bs, {x, y0}, {x + h, y0 + h});
        b.diagonal(tn, an, {x, y1}, {x - h, y1 - h});
        b.diagonal(tn, bn, {x, y1}, {x + h, y1 - h});
    }
    b.link_rings();

    // blocks as ORCA obstacles; vertex order NE, NW, SW, SE as Simulator::AddObstacleArea
    // (/root/reference/ECMAgentSimulator/Simulator.cpp:416-419), i.e. counter-clockwise.
    for (int j = 0; j < nby; j++)
        for (int i = 0; i < nbx; i++) {
            const float q[8] = {(float)bxh[i], (float)byh[j], (float)bxl[i], (float)byh[j],
                                (float)bxl[i], (float)byl[j], (float)bxh[i], (float)byl[j]};
            AppendObstacle(w.obst, q, 4);
        }
    out = std::move(w);
    return true;
}

}  // namespace ecmb200
