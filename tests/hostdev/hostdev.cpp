// TEST INFRASTRUCTURE - not part of the product, never linked into it.
//
// The per-agent device functions of csrc/device/{locate,knn,orca}.cuh compiled for the HOST (g++ with the shim
// headers in ./shim, -ffp-contract=off like nvcc's -fmad=false) and driven one agent at a time through the same
// sequence as the kernels of csrc/device/tick.cuh: counting-sort grid + snapshot, attraction point, exact
// 5-NN, ORCA, integration.  With the C library's sinf / cosf / atanf this must reproduce the reference's golden
// trajectories BIT FOR BIT (tests/test_hostdev.py): a CPU pin of the device algorithms themselves, which the GPU
// tests can only hold to 1e-4 because CUDA's sinf / cosf / atanf round differently.
//
// It mirrors, and cites, the glue it replaces: k_bin_count / k_scatter (tick.cuh), attract_agent, orca_agent,
// finish_agent, integrate_agent, and ecmgpu.cu's append_path (block bounding boxes of the polylines).
#include <vector>

static long long g_hd_considered = 0;  // neighbour candidates visited (Knn::consider), see hd_considered()
#define ECM_KNN_STATS (g_hd_considered++)

#include "../../ecmgenerator_b200/csrc/device/locate.cuh"
#include "../../ecmgenerator_b200/csrc/device/orca.cuh"

using namespace ecm;

// -DHD_SYNC: take the warp-synchronous instantiations the kernels use (kSync = true: warp-uniform trip counts,
// flags instead of early returns).  With one thread per "warp" the collectives are identities, so this checks
// the flattened control flow, not the lock-step itself.
#ifdef HD_SYNC
constexpr bool kHdSync = true;
#else
constexpr bool kHdSync = false;
#endif

namespace {

struct HdWorld {
    std::vector<float2> vert_xy, edge_cl, obst_xy, obst_dir;
    std::vector<int2> edge_v;
    std::vector<int> obst_next, obst_prev, cell_items, obst_items, cell_start, obst_start;
    std::vector<unsigned char> obst_convex;
    EcmView ecm;
    ObstView obst;
    BinView bins;
};

}  // namespace

extern "C" {

// One bin over everything: the cell list is every cell in index order, the obstacle list every segment in
// (obstacle, vertex) order - exactly the reference's linear scans (ECMCellCollection.cpp:57-90, Simulator.cpp:259-292).
void* hd_world(int nV, const float* vert_xy, int nE, const int* edge_v, const float* edge_cl, int nO, const float* obst_xy,
               const int* obst_next, const int* obst_prev, const unsigned char* obst_convex) {
    HdWorld* w = new HdWorld;
    w->vert_xy.resize(nV);
    for (int i = 0; i < nV; i++) w->vert_xy[i] = make_float2(vert_xy[2 * i], vert_xy[2 * i + 1]);
    w->edge_v.resize(nE);
    w->edge_cl.resize(4 * (size_t)nE);
    for (int i = 0; i < nE; i++) w->edge_v[i] = make_int2(edge_v[2 * i], edge_v[2 * i + 1]);
    for (int i = 0; i < 4 * nE; i++) w->edge_cl[i] = make_float2(edge_cl[2 * i], edge_cl[2 * i + 1]);
    w->obst_xy.resize(nO);
    w->obst_dir.resize(nO);
    w->obst_next.assign(obst_next, obst_next + nO);
    w->obst_prev.assign(obst_prev, obst_prev + nO);
    w->obst_convex.assign(obst_convex, obst_convex + nO);
    for (int i = 0; i < nO; i++) w->obst_xy[i] = make_float2(obst_xy[2 * i], obst_xy[2 * i + 1]);
    for (int i = 0; i < nO; i++) {  // ecmgpu_set_obstacles: Vec2::Normalize of the segment direction (ECMDataTypes.h:45-52)
        volatile float dx = obst_xy[2 * obst_next[i]] - obst_xy[2 * i], dy = obst_xy[2 * obst_next[i] + 1] - obst_xy[2 * i + 1];
        volatile float xx = dx * dx, yy = dy * dy;
        volatile float l = sqrtf(xx + yy);
        w->obst_dir[i] = l == 0.0f ? make_float2(dx, dy) : make_float2(dx / l, dy / l);
    }
    for (int c = 0; c < 2 * nE; c++) w->cell_items.push_back(c);
    for (int o = 0; o < nO; o++) w->obst_items.push_back(o);
    w->cell_start = {0, 2 * nE};
    w->obst_start = {0, nO};
    w->ecm = EcmView{nV, nE, w->vert_xy.data(), w->edge_v.data(), w->edge_cl.data()};
    w->obst = ObstView{nO, w->obst_xy.data(), w->obst_next.data(), w->obst_prev.data(), w->obst_convex.data(), w->obst_dir.data()};
    BinView b{};  // one bin over everything = the reference's linear scans (no level lists needed: n_levels = 0)
    b.x0 = -1.0e9f; b.y0 = -1.0e9f; b.inv_bin = 1.0e-12f; b.w = 1; b.h = 1;
    b.cell_start = w->cell_start.data(); b.cell_items = w->cell_items.data();
    b.obst_start = w->obst_start.data(); b.obst_items = w->obst_items.data();
    w->bins = b;
    return w;
}
void hd_world_free(void* h) { delete (HdWorld*)h; }
long long hd_considered(int reset) { long long v = g_hd_considered; if (reset) g_hd_considered = 0; return v; }

void hd_locate(void* h, int n, const float* xy, int* out_cell) {
    HdWorld* w = (HdWorld*)h;
    for (int i = 0; i < n; i++) out_cell[i] = find_cell<kHdSync>(w->ecm, w->bins, V(xy[2 * i], xy[2 * i + 1]));
}

void hd_retract(void* h, int n, const float* xy, unsigned char* ok, float* out_xy, int* out_edge) {  // k_retract
    HdWorld* w = (HdWorld*)h;
    for (int i = 0; i < n; i++) {
        v2 p = V(xy[2 * i], xy[2 * i + 1]), r = V(0.0f, 0.0f);
        int c = find_cell<kHdSync>(w->ecm, w->bins, p);
        bool good = c >= 0 && retract_in_cell(w->ecm, c, p, r);
        ok[i] = good ? 1 : 0;
        out_xy[2 * i] = r.x; out_xy[2 * i + 1] = r.y;
        out_edge[i] = c >= 0 ? (c >> 1) : -1;
    }
}

// One tick over slot arrays, in place.  cell: neighbour-grid cell size (any value must give the same result).
// path_off / path_xy: polyline of slot i = points [path_off[i], path_off[i+1]).  Events: slots whose location
// failed (replan wanted) and slots destroyed on arrival, in slot order.  nbr / nbr_cnt may be NULL.
// Returns the number of agents whose ring budget ran out (they take the exhaustive search, like k_fallback).
int hd_tick(void* h, int n, float step, float cell, int max_ring, float* pos_io, float* vel_io, float* pref_io, float* attr_io, float* force_o,
            const float* radius, const float* speed, unsigned char* active_io, const int* path_off, const float* path_xy,
            int* nbr_o, int* nbr_cnt_o, int* replan_o, int* n_replan_o, int* destroyed_o, int* n_destroyed_o, unsigned* status_o) {
    HdWorld* w = (HdWorld*)h;
    float2* pos = (float2*)pos_io; float2* vel = (float2*)vel_io; float2* pref = (float2*)pref_io; float2* attr = (float2*)attr_io;
    float2* force = (float2*)force_o;
    // ---- k_bin_count / k_scan / k_scatter: grid over the bounding box of the active agents
    float x0 = 3e38f, y0 = 3e38f, x1 = -3e38f, y1 = -3e38f;
    int n_act = 0;
    for (int i = 0; i < n; i++) if (active_io[i]) { n_act++; x0 = std::min(x0, pos[i].x); y0 = std::min(y0, pos[i].y); x1 = std::max(x1, pos[i].x); y1 = std::max(y1, pos[i].y); }
    *n_replan_o = 0; *n_destroyed_o = 0;
    if (n_act == 0) return 0;
    GridView g{};
    g.x0 = x0 - cell; g.y0 = y0 - cell; g.cell = cell; g.inv_cell = 1.0f / cell;
    g.w = (int)((x1 - g.x0) / cell) + 2; g.h = (int)((y1 - g.y0) / cell) + 2;
    std::vector<int> key(n, -1), start((size_t)g.w * g.h + 1, 0);
    for (int i = 0; i < n; i++) if (active_io[i]) { int cx, cy; g.cell_of(pos[i], cx, cy); key[i] = cy * g.w + cx; start[key[i] + 1]++; }
    for (size_t c = 1; c < start.size(); c++) start[c] += start[c - 1];
    std::vector<int> fill(start.begin(), start.end() - 1), s_slot(n_act);
    std::vector<float2> s_pos(n_act), s_vel(n_act), s_pref(n_act);
    std::vector<float> s_rad(n_act), s_spd(n_act);
    std::vector<unsigned char> s_alive(n_act, 1);
    for (int i = 0; i < n; i++) if (active_io[i]) {  // any order inside a cell is allowed: the result must not depend on it
        int p = fill[key[i]]++;
        s_slot[p] = i; s_pos[p] = pos[i]; s_vel[p] = vel[i]; s_rad[p] = radius[i]; s_spd[p] = speed[i];
    }
    g.n_sorted = n_act; g.cell_start = start.data(); g.s_pos = s_pos.data(); g.s_vel = s_vel.data(); g.s_rad = s_rad.data(); g.s_slot = s_slot.data(); g.ext_of = nullptr;
    // ---- attract_agent (UpdateAttractionPointSystem + ApplySteeringForce)
    std::vector<float2> poly;
    std::vector<float4> boxes;
    std::vector<int> replans, destroyed;
    for (int p = 0; p < n_act; p++) {
        const int slot = s_slot[p];
        unsigned st = 0u;
        const int np = path_off[slot + 1] - path_off[slot];
        poly.assign((const float2*)path_xy + path_off[slot], (const float2*)path_xy + path_off[slot + 1]);
        {   // ecmgpu.cu append_path: padded bounding boxes of blocks of 8 segments
            const int nseg = np - 1, nblk = (nseg + kPathBlock - 1) / kPathBlock;
            boxes.assign(std::max(nblk, 1), make_float4(0, 0, 0, 0));
            for (int b = 0; b < nblk; b++) {
                float4 bb = make_float4(3e38f, 3e38f, -3e38f, -3e38f);
                for (int i = b * kPathBlock; i <= std::min((b + 1) * kPathBlock, nseg); i++) {
                    bb.x = std::min(bb.x, poly[i].x); bb.y = std::min(bb.y, poly[i].y);
                    bb.z = std::max(bb.z, poly[i].x); bb.w = std::max(bb.w, poly[i].y);
                }
                const float pad = 0.05f;
                bb.x -= pad; bb.y -= pad; bb.z += pad; bb.w += pad;
                boxes[b] = bb;
            }
        }
        const v2 P = s_pos[p], goal = poly[np - 1];
        v2 a = V(0.0f, 0.0f);
        bool have = false, alive = true, need_irm = false;
        const float ddx = P.x - goal.x, ddy = P.y - goal.y;
        const float dist = ddx * ddx + ddy * ddy;
        if (dist < 20.0f * 20.0f) {
            a = goal; have = true; st |= 4u;
            if (dist < 2.0f * 2.0f) { alive = false; st |= 8u; active_io[slot] = 0; destroyed.push_back(slot); }
        } else {
            need_irm = true;
        }
        v2 ap = V(0.0f, 0.0f);
        int cellid = -2;
        const bool ok = find_attraction_point<kHdSync>(w->ecm, w->bins, P, poly.data(), boxes.data(), np, goal, ap, cellid, need_irm);
        if (need_irm) {
            if (ok) { a = ap; have = true; }
            else { st |= 2u; if (cellid == -1) st |= 1u; replans.push_back(slot); }
        }
        if (have) attr[slot] = a; else a = attr[slot];
        if (alive) {
            v2 d = vnormalized(vsub(a, P));
            v2 pv = vmul(d, s_spd[p]);
            pref[slot] = pv;
            s_pref[p] = pv;
        }
        s_alive[p] = alive ? 1 : 0;
        status_o[slot] = st;
    }
    // ---- orca_agent / finish_agent / integrate_agent
    int fallbacks = 0;
    Lp3dQueue none;
    memset(&none, 0, sizeof(none));
    for (int p = 0; p < n_act; p++) {
        if (!s_alive[p]) continue;
        const int slot = s_slot[p];
        Knn k;
        if (!knn_grid(k, s_pos[p], g, max_ring)) {  // k_fallback: exhaustive search
            fallbacks++;
            status_o[slot] |= 32u;
            k.init();
            for (int c = 0; c < n_act; c++) k.consider(s_pos[p], c, g);
        }
        const int n_nb = k.count();
        OrcaResult r = orca_velocity<kHdSync, false>(w->obst, w->bins, g, s_pos[p], s_vel[p], s_rad[p], s_spd[p], s_pref[p], n_nb, k.q, step, true, none, p);
        const v2 f = V(r.velocity.x - s_vel[p].x, r.velocity.y - s_vel[p].y);
        const float massRecip = 1.0f / 0.8f;
        const v2 nv = V(s_vel[p].x + f.x * massRecip * step, s_vel[p].y + f.y * massRecip * step);
        const v2 npos = V(s_pos[p].x + (nv.x * step), s_pos[p].y + (nv.y * step));
        force[slot] = f; vel[slot] = nv; pos[slot] = npos;
        status_o[slot] |= r.status;
        if (nbr_o) {
            for (int j = 0; j < kK; j++) nbr_o[kK * slot + j] = k.q[j] >= 0 ? s_slot[k.q[j]] : -1;
            nbr_cnt_o[slot] = n_nb;
        }
    }
    std::sort(replans.begin(), replans.end());
    std::sort(destroyed.begin(), destroyed.end());
    for (size_t i = 0; i < replans.size(); i++) replan_o[i] = replans[i];
    for (size_t i = 0; i < destroyed.size(); i++) destroyed_o[i] = destroyed[i];
    *n_replan_o = (int)replans.size();
    *n_destroyed_o = (int)destroyed.size();
    return fallbacks;
}

// Neighbour query on the current state (ecmgpu_find_neighbors): same grid build, kNN only.
void hd_neighbors(int n, float cell, int max_ring, const float* pos_in, const unsigned char* active, int* nbr_o, int* nbr_cnt_o) {
    const float2* pos = (const float2*)pos_in;
    float x0 = 3e38f, y0 = 3e38f, x1 = -3e38f, y1 = -3e38f;
    int n_act = 0;
    for (int i = 0; i < n; i++) if (active[i]) { n_act++; x0 = std::min(x0, pos[i].x); y0 = std::min(y0, pos[i].y); x1 = std::max(x1, pos[i].x); y1 = std::max(y1, pos[i].y); }
    for (int i = 0; i < n; i++) { nbr_cnt_o[i] = 0; for (int j = 0; j < kK; j++) nbr_o[kK * i + j] = -1; }
    if (n_act == 0) return;
    GridView g{};
    g.x0 = x0 - cell; g.y0 = y0 - cell; g.cell = cell; g.inv_cell = 1.0f / cell;
    g.w = (int)((x1 - g.x0) / cell) + 2; g.h = (int)((y1 - g.y0) / cell) + 2;
    std::vector<int> key(n, -1), start((size_t)g.w * g.h + 1, 0);
    for (int i = 0; i < n; i++) if (active[i]) { int cx, cy; g.cell_of(pos[i], cx, cy); key[i] = cy * g.w + cx; start[key[i] + 1]++; }
    for (size_t c = 1; c < start.size(); c++) start[c] += start[c - 1];
    std::vector<int> fill(start.begin(), start.end() - 1), s_slot(n_act);
    std::vector<float2> s_pos(n_act);
    for (int i = n - 1; i >= 0; i--) if (active[i]) { int p = fill[key[i]]++; s_slot[p] = i; s_pos[p] = pos[i]; }  // reversed on purpose
    g.n_sorted = n_act; g.cell_start = start.data(); g.s_pos = s_pos.data(); g.s_vel = nullptr; g.s_rad = nullptr; g.s_slot = s_slot.data();
    for (int p = 0; p < n_act; p++) {
        Knn k;
        if (!knn_grid(k, s_pos[p], g, max_ring)) { k.init(); for (int c = 0; c < n_act; c++) k.consider(s_pos[p], c, g); }
        const int slot = s_slot[p];
        for (int j = 0; j < kK; j++) nbr_o[kK * slot + j] = k.q[j] >= 0 ? s_slot[k.q[j]] : -1;
        nbr_cnt_o[slot] = k.count();
    }
}

}  // extern "C"
