// TEST INFRASTRUCTURE - not part of the product, never linked into it.
//
// The KERNELS of csrc/device/tick.cuh and csrc/device/strips.cuh compiled for the host and "launched" as plain
// loops with ONE thread per block (blockIdx.x walks the grid; a one-lane warp makes every warp collective an
// identity).  The sequence of launches is that of ecmgpu_update_phase (csrc/ecmgpu.cu): pack -> exchange ->
// adopt migrants -> bin count (+ ghosts) -> scan -> scatter (+ ghosts) -> k_attract -> k_orca -> k_fallback,
// with the in-process transport's message copy between strips.  What this adds to hostdev.cpp (which calls the
// device FUNCTIONS): the kernels' own glue - snapshot rows, ghosts, ownership hand-over, the LP3D queue and
// its finisher, the event lists, the record compaction of ecmgpu_update_io_owned - checked on the CPU, bit for
// bit, against the reference's golden trajectories and against the un-stripped run (tests/test_hostdev_kernels.py).
// Not emulated: the three scan kernels (a host prefix sum stands in) and the warp-per-agent half of k_fallback
// (needs 32 co-operating lanes; the scenes used never exhaust the ring budget, which the harness asserts).
#include <vector>

#include <numeric>

#include "../../ecmgenerator_b200/csrc/device/strips.cuh"
#include "../../ecmgenerator_b200/csrc/device/kdtree.cuh"
#include "../../ecmgenerator_b200/csrc/device/planner.cuh"

using namespace ecm;

namespace {

// `block`: the CTA size the product launches the kernel with.  The plain build ignores it (one thread per block);
// the -DHD_SIMT build (shim/simt.h) runs real blocks of co-operating threads.
template <class F>
void launch(int threads, F&& body, int block = 1) {
#ifdef HD_SIMT
    if (threads <= 0) return;
    hd_simt_launch((threads + block - 1) / block, block, std::function<void()>(body));
#else
    (void)block;
    gridDim.x = threads; blockDim.x = 1; threadIdx.x = 0;
    for (int b = 0; b < threads; b++) { blockIdx.x = b; body(); }
    blockIdx.x = 0; gridDim.x = 1;
#endif
}

struct Emu;
static void emu_scan_cells(Emu* e);

struct Emu {
    // world (one bin over everything, like hostdev.cpp)
    std::vector<float2> vert_xy, edge_cl, obst_xy, obst_dir;
    std::vector<int2> edge_v;
    std::vector<int> obst_next, obst_prev, cell_items, obst_items, bin_cell_start, bin_obst_start;
    std::vector<unsigned char> obst_convex;
    // slots
    int n = 0, n_slots = 0;
    float step = 0;
    std::vector<float2> pos, vel, prefvel, attraction, force;
    std::vector<float> radius, speed;
    std::vector<unsigned char> active, replan_pending;
    std::vector<unsigned> status;
    std::vector<int> cell, nbr, nbr_cnt;
    std::vector<PathHdr> path_hdr;
    std::vector<float2> path_pool;
    std::vector<float4> path_bbox;
    // grid + snapshot + scratch
    float gx0 = 0, gy0 = 0, gcell = 1;
    int gw = 1, gh = 1;
    std::vector<int> key, rank, cell_count, s_slot, fb_list, ev_replan, ev_destroyed;
    std::vector<float2> s_pos, s_vel, s_pref;
    std::vector<float> s_rad, s_spd;
    std::vector<unsigned char> s_alive, s_ghost;
    std::vector<unsigned long long> counters, scan_state;
    unsigned scan_ctl[4] = {1u, 0u, 0u, 0u};
    std::vector<int4> lp3d_hdr;
    std::vector<float4> lp3d_out, lp3d_cs;
    // strips
    bool strips = false;
    int rank_id = 0, n_ranks = 1, cap_halo = 0, cap_migr = 0, cap_self = 0;
    float lo = 0, hi = 0, halo = 0;
    std::vector<unsigned char> send[2], recv[2];
    std::vector<HaloEntry> self_ghost;
    std::vector<int> self_ghost_n, g_key, g_rank;
    // compact walk (ECMGPU_COMPACT): ecmgpu.cu ensure_walk
    bool compact = false, walk_dirty = true;
    std::vector<int> walk, walk_n;
    std::vector<unsigned char> in_walk;
    // faithful KD-tree mode (kdtree.cuh)
    std::vector<unsigned long long> kd_keys[2];
    std::vector<int> kd_vals[2], kd_seg_r[2], kd_seg_node[2], kd_raw, kd_raw_cnt, kd_cache, kd_meta;
    std::vector<float4> kd_tree;
    std::vector<float2> kd_pre_pos, kd_pre_vel;
    int kd_cap = 0;
    // device planner (planner.cuh)
    std::vector<float> vert_clear;
    std::vector<int> vert_he, he_next;

    TickView view() {
        TickView t;
        memset(&t, 0, sizeof(t));
        t.ecm = EcmView{(int)vert_xy.size(), (int)edge_v.size(), vert_xy.data(), edge_v.data(), edge_cl.data()};
        t.obst = ObstView{(int)obst_xy.size(), obst_xy.data(), obst_next.data(), obst_prev.data(), obst_convex.data(), obst_dir.data()};
        t.bins = BinView{};
        t.bins.x0 = -1.0e9f; t.bins.y0 = -1.0e9f; t.bins.inv_bin = 1.0e-12f; t.bins.w = 1; t.bins.h = 1;
        t.bins.cell_start = bin_cell_start.data(); t.bins.cell_items = cell_items.data();
        t.bins.obst_start = bin_obst_start.data(); t.bins.obst_items = obst_items.data();
        t.grid.x0 = gx0; t.grid.y0 = gy0; t.grid.cell = gcell; t.grid.inv_cell = 1.0f / gcell; t.grid.w = gw; t.grid.h = gh;
        t.grid.n_sorted = 0; t.grid.cell_start = cell_count.data();
        t.grid.s_pos = s_pos.data(); t.grid.s_vel = s_vel.data(); t.grid.s_rad = s_rad.data(); t.grid.s_slot = s_slot.data(); t.grid.ext_of = nullptr;
        t.ag.pos = pos.data(); t.ag.vel = vel.data(); t.ag.prefvel = prefvel.data(); t.ag.attraction = attraction.data();
        t.ag.force = force.data(); t.ag.radius = radius.data(); t.ag.speed = speed.data(); t.ag.active = active.data();
        t.ag.replan_pending = replan_pending.data(); t.ag.status = status.data(); t.ag.cell = cell.data();
        t.ag.nbr = nbr.data(); t.ag.nbr_cnt = nbr_cnt.data();
        t.ag.path_hdr = path_hdr.data(); t.ag.path_pool = path_pool.data(); t.ag.path_bbox = path_bbox.data();
        t.sc.key = key.data(); t.sc.rank = rank.data(); t.sc.cell_count = cell_count.data();
        t.sc.s_pos = s_pos.data(); t.sc.s_vel = s_vel.data(); t.sc.s_rad = s_rad.data(); t.sc.s_spd = s_spd.data();
        t.sc.s_slot = s_slot.data(); t.sc.s_pref = s_pref.data(); t.sc.s_alive = s_alive.data(); t.sc.s_ghost = s_ghost.data();
        t.sc.fb_list = fb_list.data(); t.sc.ev_replan = ev_replan.data(); t.sc.ev_destroyed = ev_destroyed.data(); t.sc.ev_cap = (int)ev_destroyed.size();
        t.sc.counters = counters.data();
        t.n_sorted_ptr = cell_count.data() + (size_t)gw * gh;
        t.step = step; t.max_ring = 8; t.record_neighbors = 1;
        t.strips = strips ? 1 : 0;
        const float inf = CUDART_INF_F;
        t.cover_lo = strips && rank_id > 0 ? lo - halo : -inf;
        t.cover_hi = strips && rank_id < n_ranks - 1 ? hi + halo : inf;
        t.lp3d.cap = n; t.lp3d.count = counters.data() + C_LP3D_N;
        t.lp3d.hdr = lp3d_hdr.data(); t.lp3d.out = lp3d_out.data(); t.lp3d.cs = lp3d_cs.data();
        return t;
    }
    StripView sview() {
        StripView v;
        memset(&v, 0, sizeof(v));
        v.enabled = strips ? 1 : 0; v.rank = rank_id; v.n_ranks = n_ranks; v.lo = lo; v.hi = hi; v.halo = halo;
        v.cap_halo = cap_halo; v.cap_migr = cap_migr; v.cap_self = cap_self;
        for (int d = 0; d < 2; d++) { v.send[d] = send[d].data(); v.recv[d] = recv[d].data(); v.send_hdr[d] = (MsgHeader*)send[d].data(); }
        v.self_ghost = self_ghost.data(); v.self_ghost_n = self_ghost_n.data(); v.g_key = g_key.data(); v.g_rank = g_rank.data();
        const bool w = compact && strips && !walk.empty();
        v.walk.list = w ? walk.data() : nullptr; v.walk.n = walk_n.data(); v.walk.in_list = in_walk.data();
        return v;
    }
    void append_path(int slot, const float2* pts, int np) {  // ecmgpu.cu append_path
        while (path_pool.size() % kPathBlock) path_pool.push_back(make_float2(0.0f, 0.0f));
        const int off = (int)path_pool.size();
        path_pool.insert(path_pool.end(), pts, pts + np);
        const int nseg = np - 1, nblk = (nseg + kPathBlock - 1) / kPathBlock;
        path_bbox.resize((size_t)off / kPathBlock + std::max(nblk, 1), make_float4(0, 0, 0, 0));
        const float pad = 0.05f;
        for (int b = 0; b < nblk; b++) {
            float4 bb = make_float4(3e38f, 3e38f, -3e38f, -3e38f);
            for (int i = b * kPathBlock; i <= std::min((b + 1) * kPathBlock, nseg); i++) {
                bb.x = std::min(bb.x, pts[i].x); bb.y = std::min(bb.y, pts[i].y);
                bb.z = std::max(bb.z, pts[i].x); bb.w = std::max(bb.w, pts[i].y);
            }
            bb.x -= pad; bb.y -= pad; bb.z += pad; bb.w += pad;
            path_bbox[(size_t)off / kPathBlock + b] = bb;
        }
        path_hdr[slot] = PathHdr{off, np, pts[np - 1].x, pts[np - 1].y};
    }
};

// Exclusive scan of the cell counts in place, total behind the last cell.  SIMT build: the scan kernel
// itself (tick.cuh), launched like enqueue_grid_build does; plain build: a host prefix sum stands in.
static void emu_scan_cells(Emu* e) {
#ifdef HD_SIMT
    const int tiles = (int)(e->cell_count.size() / kScanTile);
    launch(tiles * kScanBlock, [&] { k_scan_onepass((int4*)e->cell_count.data(), tiles, e->scan_state.data(), e->scan_ctl); }, kScanBlock);
#else
    int run = 0;
    for (size_t c = 0; c < (size_t)e->gw * e->gh + 1; c++) { int v = e->cell_count[c]; e->cell_count[c] = run; run += v; }
#endif
}

}  // namespace

extern "C" {

void* emu_create(int nV, const float* vert_xy, int nE, const int* edge_v, const float* edge_cl, int nO, const float* obst_xy,
                 const int* obst_next, const int* obst_prev, const unsigned char* obst_convex, int max_agents, float step,
                 float gx0, float gy0, float cell, int gw, int gh) {
    Emu* e = new Emu;
    e->vert_xy.resize(nV);
    for (int i = 0; i < nV; i++) e->vert_xy[i] = make_float2(vert_xy[2 * i], vert_xy[2 * i + 1]);
    e->edge_v.resize(nE);
    e->edge_cl.resize(4 * (size_t)nE);
    for (int i = 0; i < nE; i++) e->edge_v[i] = make_int2(edge_v[2 * i], edge_v[2 * i + 1]);
    for (int i = 0; i < 4 * nE; i++) e->edge_cl[i] = make_float2(edge_cl[2 * i], edge_cl[2 * i + 1]);
    e->obst_xy.resize(nO); e->obst_dir.resize(nO);
    e->obst_next.assign(obst_next, obst_next + nO); e->obst_prev.assign(obst_prev, obst_prev + nO);
    e->obst_convex.assign(obst_convex, obst_convex + nO);
    for (int i = 0; i < nO; i++) e->obst_xy[i] = make_float2(obst_xy[2 * i], obst_xy[2 * i + 1]);
    for (int i = 0; i < nO; i++) {
        volatile float dx = obst_xy[2 * obst_next[i]] - obst_xy[2 * i], dy = obst_xy[2 * obst_next[i] + 1] - obst_xy[2 * i + 1];
        volatile float xx = dx * dx, yy = dy * dy;
        volatile float l = sqrtf(xx + yy);
        e->obst_dir[i] = l == 0.0f ? make_float2(dx, dy) : make_float2(dx / l, dy / l);
    }
    for (int c = 0; c < 2 * nE; c++) e->cell_items.push_back(c);
    for (int o = 0; o < nO; o++) e->obst_items.push_back(o);
    e->bin_cell_start = {0, 2 * nE};
    e->bin_obst_start = {0, nO};
    const int n = max_agents;
    e->n = n; e->step = step; e->gx0 = gx0; e->gy0 = gy0; e->gcell = cell; e->gw = gw; e->gh = gh;
    e->pos.assign(n, make_float2(0, 0)); e->vel = e->prefvel = e->attraction = e->force = e->pos;
    e->radius.assign(n, 0); e->speed.assign(n, 0); e->active.assign(n, 0); e->replan_pending.assign(n, 0);
    e->status.assign(n, 0); e->cell.assign(n, -2); e->nbr.assign(5 * (size_t)n, -1); e->nbr_cnt.assign(n, 0);
    e->path_hdr.assign(n, PathHdr{0, 0, 0.0f, 0.0f});
    e->key.assign(n, -1); e->rank.assign(n, 0);
    const size_t padded = (((size_t)gw * gh + 1 + kScanTile - 1) / kScanTile) * kScanTile;  // ecmgpu.cu build_grid: ncells_padded
    e->cell_count.assign(padded, 0); e->scan_state.assign(padded / kScanTile, 0ull);
    e->counters.assign(C_COUNT, 0ull);
    e->ev_replan.assign(n, 0); e->ev_destroyed.assign(n, 0);
    e->lp3d_hdr.resize(n); e->lp3d_out.resize(n); e->lp3d_cs.resize((size_t)n * kMaxCons);
    const size_t cap = n;  // grows in emu_set_strips
    e->s_pos.resize(cap); e->s_vel.resize(cap); e->s_pref.resize(cap); e->s_rad.resize(cap); e->s_spd.resize(cap);
    e->s_slot.resize(cap); e->s_alive.resize(cap); e->s_ghost.assign(cap, 0); e->fb_list.resize(cap);
    return e;
}
void emu_destroy(void* h) { delete (Emu*)h; }
void emu_set_compact(void* h, int on) { Emu* e = (Emu*)h; e->compact = on != 0; e->walk_dirty = true; }
int emu_walk_len(void* h) { Emu* e = (Emu*)h; return e->walk_n.empty() ? -1 : e->walk_n[0]; }
// ecmgpu.cu ensure_walk
static void emu_ensure_walk(Emu* e) {
    if (!e->compact || !e->strips) return;
    if (e->walk.empty()) { e->walk.assign(e->n, 0); e->walk_n.assign(1, 0); e->in_walk.assign(e->n, 0); e->walk_dirty = true; }
    if (!e->walk_dirty) return;
    e->walk_n[0] = 0;
    std::fill(e->in_walk.begin(), e->in_walk.end(), 0);
    WalkView w{e->walk.data(), e->walk_n.data(), e->in_walk.data()};
    launch(e->n_slots, [&] { k_walk_rebuild(e->n_slots, e->active.data(), w); }, kPackBlock);
    e->walk_dirty = false;
}

void emu_load(void* h, int n, const float* pos, const float* radius, const float* speed, const int* path_off, const float* path_xy) {
    Emu* e = (Emu*)h;
    for (int i = 0; i < n; i++) {
        e->pos[i] = make_float2(pos[2 * i], pos[2 * i + 1]);
        e->vel[i] = e->prefvel[i] = e->attraction[i] = e->force[i] = make_float2(0, 0);
        e->radius[i] = radius[i]; e->speed[i] = speed[i]; e->active[i] = 1; e->replan_pending[i] = 0;
        e->append_path(i, (const float2*)path_xy + path_off[i], path_off[i + 1] - path_off[i]);
    }
    e->n_slots = std::max(e->n_slots, n);
    e->walk_dirty = true;
}
void emu_set_path(void* h, int slot, const float* xy, int np) {
    Emu* e = (Emu*)h;
    e->append_path(slot, (const float2*)xy, np);
    e->replan_pending[slot] = 0;
}
void emu_destroy_agent(void* h, int slot) { ((Emu*)h)->active[slot] = 0; }
// ecmgpu_write(VEL / ATTRACTION): a mid-run state for single-tick replays
void emu_set_state(void* h, const float* vel, const float* attraction) {
    Emu* e = (Emu*)h;
    for (int i = 0; i < e->n_slots; i++) {
        e->vel[i] = make_float2(vel[2 * i], vel[2 * i + 1]);
        e->attraction[i] = make_float2(attraction[2 * i], attraction[2 * i + 1]);
    }
}

// ecmgpu_comm_set_strips (fixed capacities given by the caller) + k_assign_owner
void emu_set_strips(void* h, int rank, int n_ranks, float lo, float hi, float halo, int cap_halo, int cap_migr) {
    Emu* e = (Emu*)h;
    e->strips = true; e->rank_id = rank; e->n_ranks = n_ranks; e->lo = lo; e->hi = hi; e->halo = halo;
    e->cap_halo = cap_halo; e->cap_migr = cap_migr; e->cap_self = 2 * cap_migr;
    const size_t msg = strip_msg_bytes(cap_halo, cap_migr);
    for (int d = 0; d < 2; d++) { e->send[d].assign(msg, 0); e->recv[d].assign(msg, 0); }
    e->self_ghost.resize(e->cap_self); e->self_ghost_n.assign(1, 0);
    const size_t ng = 2 * (size_t)cap_halo + e->cap_self;
    e->g_key.assign(ng, -1); e->g_rank.assign(ng, 0);
    const size_t cap = e->n + ng;
    e->s_pos.resize(cap); e->s_vel.resize(cap); e->s_pref.resize(cap); e->s_rad.resize(cap); e->s_spd.resize(cap);
    e->s_slot.resize(cap); e->s_alive.resize(cap); e->s_ghost.assign(cap, 0); e->fb_list.resize(cap);
    e->walk_dirty = true;
    TickView t = e->view();
    StripView sv = e->sview();
    launch(e->n_slots, [&] { k_assign_owner(e->n_slots, t.ag, sv); }, 256);
}

// phase 0: enqueue_pack
void emu_pack(void* h) {
    Emu* e = (Emu*)h;
    if (!e->strips) return;
    for (int d = 0; d < 2; d++) memset(e->send[d].data(), 0, sizeof(MsgHeader));
    e->self_ghost_n[0] = 0;
    emu_ensure_walk(e);
    TickView t = e->view();
    StripView sv = e->sview();
    if (sv.walk.list) launch(37, [&] { k_pack_walk(t.ag, sv, e->counters.data()); }, kPackBlock);  // a fixed grid, several trips
    else launch(e->n_slots, [&] { k_pack(e->n_slots, t.ag, sv, e->counters.data()); }, kPackBlock);
}
// phase 1: enqueue_exchange with the in-process transport: my left neighbour's RIGHT message is my left inbox
void emu_exchange(void* h, void* left, void* right) {
    Emu* e = (Emu*)h;
    if (!e->strips) return;
    if (left) e->recv[0] = ((Emu*)left)->send[1];
    if (right) e->recv[1] = ((Emu*)right)->send[0];
    TickView t = e->view();
    StripView sv = e->sview();
    launch(std::max(e->cap_migr, 1), [&] { k_unpack_migrants(t.ag, sv); }, 256);
}
// The same exchange for strips that live in DIFFERENT processes (tests: torch.distributed / gloo carries the bytes
// the way NCCL send/recv does on the GPUs): the outgoing message of direction d, the incoming one, then adopt.
int emu_msg_bytes(void* h) { Emu* e = (Emu*)h; return (int)strip_msg_bytes(e->cap_halo, e->cap_migr); }
void emu_get_send(void* h, int d, unsigned char* out) { Emu* e = (Emu*)h; memcpy(out, e->send[d].data(), e->send[d].size()); }
void emu_set_recv(void* h, int d, const unsigned char* in) { Emu* e = (Emu*)h; memcpy(e->recv[d].data(), in, e->recv[d].size()); }
void emu_adopt(void* h) {
    Emu* e = (Emu*)h;
    TickView t = e->view();
    StripView sv = e->sview();
    launch(std::max(e->cap_migr, 1), [&] { k_unpack_migrants(t.ag, sv); }, 256);
}
// phase 2: enqueue_grid_build + k_attract + k_orca + k_fallback.  Returns the number of ring-budget fallbacks (must be 0).
int emu_tick(void* h) {
    Emu* e = (Emu*)h;
    TickView t = e->view();
    StripView sv = e->sview();
    GridParams gp{e->gx0, e->gy0, e->gcell, 1.0f / e->gcell, e->gw, e->gh};
    std::fill(e->cell_count.begin(), e->cell_count.end(), 0);
    e->counters[C_FALLBACK_N] = 0; e->counters[C_LP3D_N] = 0;
    if (sv.walk.list) launch(53, [&] { k_bin_count_walk(sv, e->active.data(), e->pos.data(), gp, e->cell_count.data(), e->key.data(), e->rank.data(), e->status.data(), e->counters.data()); }, 256);
    else launch(e->n_slots, [&] { k_bin_count(e->n_slots, e->active.data(), e->pos.data(), gp, e->cell_count.data(), e->key.data(), e->rank.data(), e->status.data(), e->counters.data()); }, 256);
    const int ng = 2 * e->cap_halo + e->cap_self;
    if (e->strips && !sv.walk.list) launch(ng, [&] { k_ghost_count(sv, gp, e->cell_count.data()); }, 256);
    emu_scan_cells(e);
    if (sv.walk.list) launch(53, [&] { k_scatter_walk(sv, e->key.data(), e->rank.data(), e->cell_count.data(), t.ag, t.sc); }, 256);
    else launch(e->n_slots, [&] { k_scatter(e->n_slots, e->key.data(), e->rank.data(), e->cell_count.data(), t.ag, t.sc); }, 256);
    if (e->strips && !sv.walk.list) launch(ng, [&] { k_ghost_scatter(sv, e->cell_count.data(), t.ag, t.sc); }, 256);
    const int rows = e->n_slots + (e->strips ? ng : 0);
    if (sv.walk.list) {  // ecmgpu_update_phase: compact strips run the fixed-grid versions
        launch(41, [&] { k_attract_tiles(t); }, 128);
        launch(29, [&] { k_orca_tiles(t); }, 256);
    } else {
        launch(rows, [&] { k_attract(t); }, 128);
        launch(rows, [&] { k_orca(t); }, 256);
    }
    const int fb = (int)e->counters[C_FALLBACK_N];
#ifdef HD_SIMT
    launch(4 * 128, [&] { k_fallback(t, 0); }, 128);  // real warps: the exhaustive warp-per-agent search too
#else
    if (fb == 0) launch(64, [&] { k_fallback(t, 0); }, 128);  // the parked LP3D agents (its warp-per-agent half needs real warps)
#endif
    return fb;
}

// ---- the tick in KD-tree neighbour mode: enqueue_grid_build + k_attract + enqueue_kd_orca + k_fallback (ecmgpu.cu).
// The library radix sort of every tree level is stood in for by std::stable_sort on the same keys.
static int kd_levels(int n) { int L = 0; while (((1ll << L) - 1) < (long long)n) L++; return L; }

static void emu_kd_alloc(Emu* e) {
    if (!e->kd_raw.empty()) return;
    const size_t n = e->n;
    for (int b = 0; b < 2; b++) { e->kd_keys[b].assign(n, 0); e->kd_vals[b].assign(n, -1); e->kd_seg_r[b].assign(n + 1, 0); e->kd_seg_node[b].assign(n + 1, 0); }
    e->kd_cap = (int)((1ll << kd_levels((int)n)) - 1);
    e->kd_tree.resize(e->kd_cap);
    e->kd_raw.assign(5 * n, -1); e->kd_raw_cnt.assign(n, 0); e->kd_cache.assign(10, 0); e->kd_meta.assign(4, 0);
    e->kd_pre_pos.resize(n); e->kd_pre_vel.resize(n);
}

static void emu_kd_build(Emu* e, const TickView& t) {
    const int n = e->n_slots;
    KdBuild b;
    b.n_slots = n; b.n_active_ptr = t.n_sorted_ptr; b.cell_key = e->key.data(); b.pos = e->pos.data();
    b.tree = e->kd_tree.data(); b.cap = e->kd_cap; b.meta = e->kd_meta.data(); b.ties = e->counters.data() + C_TOTAL_KD_TIES;
    b.small_ties = e->counters.data() + C_TOTAL_KD_SMALL_TIES;
    memset(e->kd_tree.data(), 0xff, sizeof(float4) * e->kd_tree.size());
    int in = 0;
    launch(n, [&] { k_kd_init(b, e->kd_keys[0].data(), e->kd_vals[0].data(), e->kd_seg_r[0].data(), e->kd_seg_node[0].data()); }, 256);
    const int levels = kd_levels(n);
    std::vector<int> perm(n);
    for (int d = 0; d < levels; d++) {
        std::iota(perm.begin(), perm.end(), 0);
        const unsigned long long* k = e->kd_keys[in].data();
        std::stable_sort(perm.begin(), perm.end(), [k](int a, int c) { return k[a] < k[c]; });
        std::vector<unsigned long long> sk(n);
        std::vector<int> sv(n);
        for (int i = 0; i < n; i++) { sk[i] = k[perm[i]]; sv[i] = e->kd_vals[in][perm[i]]; }
        const int out = in ^ 1;
        launch(n, [&] {
            k_kd_split(b, d, sk.data(), sv.data(), e->kd_keys[out].data(), e->kd_vals[out].data(), e->kd_seg_r[d & 1].data(), e->kd_seg_node[d & 1].data(),
                       e->kd_seg_r[(d + 1) & 1].data(), e->kd_seg_node[(d + 1) & 1].data());
        }, 256);
        in = out;
    }
}

static KdQuery emu_kd_query(Emu* e, bool carried) {
    KdQuery q;
    q.tree = e->kd_tree.data(); q.cap = e->kd_cap; q.meta = e->kd_meta.data(); q.raw = e->kd_raw.data(); q.raw_cnt = e->kd_raw_cnt.data();
    q.cache = e->kd_cache.data() + (carried ? 0 : 5);
    return q;
}

void emu_kd_reset(void* h) {  // ecmgpu_set_neighbor_mode(KDTREE)
    Emu* e = (Emu*)h;
    emu_kd_alloc(e);
    std::fill(e->kd_cache.begin(), e->kd_cache.end(), 0);
}

static void emu_grid_build(Emu* e, const TickView& t) {
    GridParams gp{e->gx0, e->gy0, e->gcell, 1.0f / e->gcell, e->gw, e->gh};
    std::fill(e->cell_count.begin(), e->cell_count.end(), 0);
    e->counters[C_FALLBACK_N] = 0; e->counters[C_LP3D_N] = 0;
    launch(e->n_slots, [&] { k_bin_count(e->n_slots, e->active.data(), e->pos.data(), gp, e->cell_count.data(), e->key.data(), e->rank.data(), e->status.data(), e->counters.data()); }, 256);
    emu_scan_cells(e);
    launch(e->n_slots, [&] { k_scatter(e->n_slots, e->key.data(), e->rank.data(), e->cell_count.data(), t.ag, t.sc); }, 256);
}

void emu_tick_kd(void* h) {
    Emu* e = (Emu*)h;
    emu_kd_alloc(e);
    TickView t = e->view();
    emu_grid_build(e, t);
    const int rows = e->n_slots;
    launch(rows, [&] { k_attract(t); }, 128);
    emu_kd_build(e, t);
    std::copy(e->pos.begin(), e->pos.begin() + e->n_slots, e->kd_pre_pos.begin());
    std::copy(e->vel.begin(), e->vel.begin() + e->n_slots, e->kd_pre_vel.begin());
    const KdQuery q = emu_kd_query(e, true);
    launch(rows, [&] { k_kd_query(t, q, 1); }, 128);
    launch(e->n_slots, [&] { k_kd_resolve(e->n_slots, e->active.data(), q, e->nbr.data(), e->nbr_cnt.data()); }, 256);
    launch(1, [&] { k_kd_cache(e->n_slots, e->active.data(), e->nbr.data(), q.cache); }, 256);
    TickView t2 = t;
    t2.grid.s_pos = e->kd_pre_pos.data(); t2.grid.s_vel = e->kd_pre_vel.data(); t2.grid.s_rad = e->radius.data();
    t2.record_neighbors = 0;
    launch(rows, [&] { k_orca_kd(t2); }, 256);
    launch(64, [&] { k_fallback(t, 0); }, 128);
}

// k_kd_resolve + k_kd_cache on hand-made search results (ids or tokens -2 - place), for tests of the token chains
void emu_kd_resolve(int n_slots, const unsigned char* active, const int* raw, const int* raw_cnt, int* cache, int* nbr, int* nbr_cnt) {
    KdQuery q;
    memset(&q, 0, sizeof(q));
    q.raw = (int*)raw; q.raw_cnt = (int*)raw_cnt; q.cache = cache;
    launch(n_slots, [&] { k_kd_resolve(n_slots, active, q, nbr, nbr_cnt); }, 256);
    launch(1, [&] { k_kd_cache(n_slots, active, nbr, cache); }, 256);
}

// ecmgpu_find_neighbors in KD-tree mode
void emu_query_neighbors_kd(void* h, int* ids, int* cnt) {
    Emu* e = (Emu*)h;
    emu_kd_alloc(e);
    TickView t = e->view();
    emu_grid_build(e, t);
    emu_kd_build(e, t);
    std::fill(e->nbr.begin(), e->nbr.end(), -1);
    std::fill(e->nbr_cnt.begin(), e->nbr_cnt.end(), -1);
    const KdQuery q = emu_kd_query(e, false);
    std::fill(q.cache, q.cache + 5, 0);
    launch(e->n_slots, [&] { k_kd_query(t, q, 0); }, 128);
    launch(e->n_slots, [&] { k_kd_resolve(e->n_slots, e->active.data(), q, e->nbr.data(), e->nbr_cnt.data()); }, 256);
    memcpy(ids, e->nbr.data(), 20 * (size_t)e->n);
    memcpy(cnt, e->nbr_cnt.data(), 4 * (size_t)e->n);
}

// ecmgpu_set_ecm_topology + ecmgpu_plan_paths with `workers` workers (queries in a grid-stride loop, like the kernel)
void emu_set_topology(void* h, const float* vert_clear, const int* vert_he, const int* he_next) {
    Emu* e = (Emu*)h;
    e->vert_clear.assign(vert_clear, vert_clear + e->vert_xy.size());
    e->vert_he.assign(vert_he, vert_he + e->vert_xy.size());
    e->he_next.assign(he_next, he_next + 2 * e->edge_v.size());
}
int emu_plan_paths(void* h, int workers, int n, const float* start, const float* goal, const float* clearance, int* out_off, int* out_len,
                   unsigned char* out_status, float* pool, int pool_cap, int cap_path, int cap_portals, int cap_out, int cap_push) {
    Emu* e = (Emu*)h;
    TickView t = e->view();
    PlanView w;
    w.ecm = t.ecm; w.bins = t.bins;
    w.vert_clear = e->vert_clear.data(); w.vert_he = e->vert_he.data(); w.he_next = e->he_next.data();
    const int nV = (int)e->vert_xy.size(), nE = (int)e->edge_v.size();
    PlanScratch sc;
    // cap_push <= 0: room for every possible push (2E + 4); else the first-pass capacity of ecmgpu_plan_paths
    sc.n_workers = workers; sc.cap_push = cap_push > 0 ? std::min(cap_push, 2 * nE + 4) : 2 * nE + 4; sc.cap_path = cap_path; sc.cap_portals = cap_portals; sc.cap_out = cap_out;
    std::vector<PlanNode> node((size_t)workers * nV);
    std::vector<int> heap((size_t)workers * sc.cap_push), touched((size_t)workers * sc.cap_push),
        vpath((size_t)workers * cap_path), epath((size_t)workers * cap_path);
    std::vector<float4> portals((size_t)workers * cap_portals);
    std::vector<float2> out((size_t)workers * cap_out);
    sc.node = node.data(); sc.heap = heap.data(); sc.touched = touched.data();
    sc.vpath = vpath.data(); sc.epath = epath.data(); sc.portals = portals.data(); sc.out = out.data();
    launch(64, [&] { k_plan_init(sc, nV); }, 256);
    int cursor = 0;
    launch(workers, [&] {
        k_plan_paths(w, sc, n, nullptr, (const float2*)start, (const float2*)goal, clearance, out_off, out_len, out_status, (float2*)pool, pool_cap, &cursor);
    }, 128);
    // every query must leave the A* records idle again (CleanRequestData through the touched list)
    for (size_t i = 0; i < node.size(); i++)
        if (node[i].g != kMaxFloat || node[i].f != kMaxFloat || node[i].parent != nV || node[i].visited) return -1;
    return cursor;
}

// ecmgpu_valid_spawn_locations
void emu_valid_spawn(void* h, int n, const float* xy, const float* clearance, unsigned char* out) {
    Emu* e = (Emu*)h;
    TickView t = e->view();
    emu_grid_build(e, t);
    launch(n, [&] { k_valid_spawn(t.grid, n, (const float2*)xy, clearance, out); }, 128);
}

void emu_read(void* h, float* pos, float* vel, float* pref, float* attr, float* force, unsigned char* active, int* nbr, int* nbr_cnt,
              unsigned* status, int* cell) {
    Emu* e = (Emu*)h;
    const size_t n = e->n;
    if (pos) memcpy(pos, e->pos.data(), 8 * n);
    if (vel) memcpy(vel, e->vel.data(), 8 * n);
    if (pref) memcpy(pref, e->prefvel.data(), 8 * n);
    if (attr) memcpy(attr, e->attraction.data(), 8 * n);
    if (force) memcpy(force, e->force.data(), 8 * n);
    if (active) memcpy(active, e->active.data(), n);
    if (nbr) memcpy(nbr, e->nbr.data(), 20 * n);
    if (nbr_cnt) memcpy(nbr_cnt, e->nbr_cnt.data(), 4 * n);
    if (status) memcpy(status, e->status.data(), 4 * n);
    if (cell) memcpy(cell, e->cell.data(), 4 * n);
}
// counters: out[0..C_COUNT)
void emu_counters(void* h, unsigned long long* out) { memcpy(out, ((Emu*)h)->counters.data(), sizeof(unsigned long long) * C_COUNT); }
// ecmgpu_poll_events: replans / destroyed since the last poll, ascending slots
void emu_poll(void* h, int* replans, int* n_replans, int* destroyed, int* n_destroyed) {
    Emu* e = (Emu*)h;
    const int nr = (int)e->counters[C_REPLAN_N], nd = (int)e->counters[C_DESTROYED_N];
    std::copy(e->ev_replan.begin(), e->ev_replan.begin() + nr, replans);
    std::copy(e->ev_destroyed.begin(), e->ev_destroyed.begin() + nd, destroyed);
    std::sort(replans, replans + nr);
    std::sort(destroyed, destroyed + nd);
    *n_replans = nr; *n_destroyed = nd;
    e->counters[C_REPLAN_N] = 0; e->counters[C_DESTROYED_N] = 0;
}
// ecmgpu_update_io_owned's two kernels
int emu_collect_owned(void* h, AgentRec* out) {
    Emu* e = (Emu*)h;
    int count = 0;
    StripView sv = e->sview();
    if (sv.walk.list && !e->walk_dirty) launch(19, [&] { k_collect_owned_walk(sv.walk, e->active.data(), e->pos.data(), e->vel.data(), out, e->n, &count, nullptr); }, kCollectBlock);
    else launch(e->n_slots, [&] { k_collect_owned(e->n_slots, e->active.data(), e->pos.data(), e->vel.data(), out, e->n, &count, nullptr); }, kCollectBlock);
    return count;
}
void emu_apply_records(void* h, int n, const AgentRec* rec) {
    Emu* e = (Emu*)h;
    launch(n, [&] { k_apply_records(n, rec, e->n, e->active.data(), e->pos.data(), e->vel.data(), nullptr); }, 256);
}

}  // extern "C"
