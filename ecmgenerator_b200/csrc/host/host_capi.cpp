// C ABI of the host-side helpers (include/ecm_b200_host.h).
#include "../../../include/ecm_b200_host.h"

#include <algorithm>
#include <cstring>
#include <string>

#include "flat_world.h"
#include "lattice_world.h"
#include "polygon_world.h"
#include "planner.h"

struct ecmhost_world {
    ecmb200::FlatWorld w;
};

struct ecmhost_paths {
    int n = 0, ok = 0;
    std::vector<int> off;
    std::vector<float> xy;
};

extern "C" {

ecmhost_world* ecmhost_lattice_world(int nbx, const float* bx, int nby, const float* by, float W, float x0,
                                     float y0) {
    auto* h = new ecmhost_world();
    if (!bx || !by || !ecmb200::BuildLatticeWorld(nbx, bx, nby, by, W, x0, y0, h->w)) {
        delete h;
        return nullptr;
    }
    return h;
}

ecmhost_world* ecmhost_polygon_world(const float bbox[4], int n_polys, const int* poly_first, const float* poly_xy, char* error, int error_cap) {
    auto* h = new ecmhost_world();
    std::string msg;
    if (!bbox || !ecmb200::BuildPolygonWorld(bbox, n_polys, poly_first, poly_xy, h->w, &msg)) {
        if (error && error_cap > 0) {
            const size_t m = std::min<size_t>(msg.size(), (size_t)error_cap - 1);
            memcpy(error, msg.data(), m);
            error[m] = 0;
        }
        delete h;
        return nullptr;
    }
    return h;
}

ecmhost_world* ecmhost_world_from_arrays(const ecmhost_world_view* v) {
    if (!v) return nullptr;
    auto* h = new ecmhost_world();
    auto& w = h->w;
    for (int i = 0; i < 4; i++) w.bbox[i] = v->bbox[i];
    w.ecm.vert_xy.assign(v->vert_xy, v->vert_xy + 2 * v->n_vertices);
    w.ecm.vert_clear.assign(v->vert_clear, v->vert_clear + v->n_vertices);
    w.ecm.vert_he.assign(v->vert_he, v->vert_he + v->n_vertices);
    w.ecm.edge_v.assign(v->edge_v, v->edge_v + 2 * v->n_edges);
    w.ecm.edge_cl.assign(v->edge_cl, v->edge_cl + 8 * v->n_edges);
    w.ecm.he_next.assign(v->he_next, v->he_next + 2 * v->n_edges);
    w.obst.xy.assign(v->obst_xy, v->obst_xy + 2 * v->n_obst_vertices);
    w.obst.next.assign(v->obst_next, v->obst_next + v->n_obst_vertices);
    w.obst.prev.assign(v->obst_prev, v->obst_prev + v->n_obst_vertices);
    w.obst.convex.assign(v->obst_convex, v->obst_convex + v->n_obst_vertices);
    w.obst.first.assign(v->obst_first, v->obst_first + v->n_obstacles + 1);
    return h;
}

void ecmhost_world_free(ecmhost_world* w) { delete w; }

int ecmhost_world_get_view(const ecmhost_world* h, ecmhost_world_view* o) {
    if (!h || !o) return -1;
    const auto& w = h->w;
    for (int i = 0; i < 4; i++) o->bbox[i] = w.bbox[i];
    o->n_vertices = w.ecm.num_vertices();
    o->n_edges = w.ecm.num_edges();
    o->n_obst_vertices = w.obst.num_vertices();
    o->n_obstacles = w.obst.num_obstacles();
    o->vert_xy = w.ecm.vert_xy.data();
    o->vert_clear = w.ecm.vert_clear.data();
    o->vert_he = w.ecm.vert_he.data();
    o->edge_v = w.ecm.edge_v.data();
    o->edge_cl = w.ecm.edge_cl.data();
    o->he_next = w.ecm.he_next.data();
    o->obst_xy = w.obst.xy.data();
    o->obst_next = w.obst.next.data();
    o->obst_prev = w.obst.prev.data();
    o->obst_convex = w.obst.convex.data();
    o->obst_first = w.obst.first.data();
    return 0;
}

ecmhost_paths* ecmhost_plan_paths(const ecmhost_world* w, int n, const float* start_xy, const float* goal_xy,
                                  const float* clearance, int threads) {
    if (!w || n < 0 || (n > 0 && (!start_xy || !goal_xy || !clearance))) return nullptr;
    auto* p = new ecmhost_paths();
    p->n = n;
    p->ok = ecmb200::PlanPaths(w->w, n, start_xy, goal_xy, clearance, threads, p->off, p->xy);
    return p;
}
int ecmhost_paths_count(const ecmhost_paths* p) { return p ? p->n : 0; }
int ecmhost_paths_succeeded(const ecmhost_paths* p) { return p ? p->ok : 0; }
const int* ecmhost_paths_offsets(const ecmhost_paths* p) { return p ? p->off.data() : nullptr; }
const float* ecmhost_paths_xy(const ecmhost_paths* p) { return p ? p->xy.data() : nullptr; }
void ecmhost_paths_free(ecmhost_paths* p) { delete p; }

int ecmhost_find_cells(const ecmhost_world* w, int n, const float* xy, int* out_cell) {
    if (!w || n < 0 || (n > 0 && (!xy || !out_cell))) return -1;
    ecmb200::CellLocator loc;
    loc.Build(w->w);
    for (int i = 0; i < n; i++) out_cell[i] = loc.FindCell(w->w, xy[2 * i], xy[2 * i + 1]);
    return 0;
}

}  // extern "C"
