// libecmgpu.so - implementation of the C ABI in include/ecm_b200.h.
//
// Owns all device state of one simulator: the flattened static world (ECM, obstacles, static bins),
// the per-slot agent components, the path pool, the per-tick neighbour grid + snapshot, and the
// event queues.  One CUDA stream per handle; ecmgpu_update() enqueues the kernels of device/tick.cuh.
// No CPU fallback: every compute entry point needs a CUDA device.
#include "../../include/ecm_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include <dlfcn.h>

#include "host/flat_world.h"
#include "device/strips.cuh"
#include "device/kdtree.cuh"
#include "device/planner.cuh"

#include <cub/device/device_radix_sort.cuh>

using namespace ecm;

namespace {

thread_local std::string g_create_error;

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;  // owns its allocation: temporaries release it on every return path
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { free(); }
    cudaError_t alloc(size_t count) {
        free();
        n = count;
        if (count == 0) return cudaSuccess;
        return cudaMalloc((void**)&p, count * sizeof(T));
    }
    void free() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
};

}  // namespace

struct ecmgpu_sim {
    ecmgpu_params prm{};
    cudaStream_t stream = nullptr;
    std::string err;

    // ---- static world, host copies (bins are (re)built from these)
    float bbox[4] = {0, 0, 0, 0};
    std::vector<float> h_vert_xy, h_edge_cl, h_obst_xy;
    std::vector<int> h_edge_v, h_obst_next;
    bool have_ecm = false, have_obst = false, bins_dirty = true;
    float built_range = 0.0f;   // obstacle range the bins were built for
    float tracked_range = 0.0f; // max over loaded agents of 10*speed + radius
    float static_bin = 0.0f;
    int bins_w = 0, bins_h = 0, max_cell_list = 0, max_obst_list = 0;
    float bins_x0 = 0, bins_y0 = 0;

    // ---- static world, device
    DevBuf<float2> d_vert_xy, d_edge_cl, d_obst_xy;
    DevBuf<int2> d_edge_v;
    DevBuf<int> d_obst_next, d_obst_prev;
    DevBuf<unsigned char> d_obst_convex;
    DevBuf<float2> d_obst_dir;
    DevBuf<int> d_bin_cell_start, d_bin_cell_items, d_bin_obst_start, d_bin_obst_items, d_row_start, d_row_items;
    DevBuf<float> d_level_y;
    DevBuf<unsigned> d_level_bits;
    int n_levels = 0, level_shift = 31;
    bool bins_closed = false;
    int n_vertices = 0, n_edges = 0, n_obst = 0;

    // ---- agents, device (per slot)
    DevBuf<float2> d_pos, d_vel, d_prefvel, d_attraction, d_force;
    DevBuf<float> d_radius, d_speed;
    DevBuf<unsigned char> d_active, d_replan_pending;
    DevBuf<unsigned> d_status;
    DevBuf<int> d_cell, d_nbr, d_nbr_cnt;
    DevBuf<PathHdr> d_path_hdr;
    DevBuf<float2> d_path_pool;
    DevBuf<float4> d_path_bbox;
    // host mirror of the path pool (append-only, compacted on overflow)
    std::vector<PathHdr> h_path_hdr;
    std::vector<float2> h_path_pool;  // every path starts on a multiple of 8 points
    std::vector<float4> h_path_bbox;  // [h_path_pool.size()/8 rounded up]
    size_t pool_uploaded = 0;  // prefix of h_path_pool already resident on the device
    int n_slots = 0;

    // ---- spatial renumbering.  Every per-agent device array is indexed by an INTERNAL index; the caller's slot id maps
    // to it through int_of / ext_of.  ecmgpu_bulk_load hands the internal indices of the slots it loads out in spatial
    // order (row-major 4 m cells), so agents that are close in the world are close in memory: the snapshot scatter, the
    // path-header gathers and the per-agent result stores of a warp then touch a few sectors instead of 32 (round 1
    // measured 3.2x DRAM write amplification in k_orca and 75 MB of pure overhead in k_scatter with slots in load
    // order).  Identity until the first large load; the KD-tree mode (which walks slots in ascending order) restores it.
    bool coherent = true;  // env ECMGPU_COHERENT=0 disables
    bool perm_identity = true;
    std::vector<int> h_int_of, h_ext_of;
    DevBuf<int> d_int_of, d_ext_of;
    DevBuf<unsigned char> d_xfer;  // staging of ecmgpu_read / ecmgpu_write while the mapping is not the identity

    // ---- neighbour grid + snapshot
    float cell = 0.0f;
    int gw = 0, gh = 0, ncells_padded = 0;
    float gx0 = 0, gy0 = 0;
    bool grid_dirty = true;
    int grid_n_slots = 0;  // n_slots when the (automatic) cell was last chosen
    DevBuf<int> d_key, d_rank, d_cell_count, d_s_slot, d_fb_list, d_ev_replan, d_ev_destroyed;
    // LP3D queue (orca.cuh Lp3dQueue): one row per slot, 32 + 16 * kMaxCons bytes each
    DevBuf<int4> d_lp3d_hdr;
    DevBuf<float4> d_lp3d_out, d_lp3d_cs;
    int lp3d_cap = 0;
    DevBuf<float2> d_s_pos, d_s_vel, d_s_pref;
    DevBuf<float> d_s_rad, d_s_spd;
    DevBuf<unsigned char> d_s_alive;
    DevBuf<unsigned long long> d_counters, d_scan_state;  // scan: one look-back word per tile of 4096 cells
    DevBuf<unsigned> d_scan_ctl;                          // scan: epoch, finished tiles, next ticket

    // ---- multi-GPU strips (device/strips.cuh)
    bool strips_on = false;
    int rank = 0, n_ranks = 1;
    float strip_lo = 0, strip_hi = 0, halo = 0;
    int cap_halo = 0, cap_migr = 0, cap_self = 0;
    void* nccl_comm = nullptr;
    ecmgpu_sim* peer[2] = {nullptr, nullptr};  // in-process transport: left / right neighbour handles
    bool local_transport = false;
    cudaEvent_t ev_packed = nullptr, ev_pulled = nullptr;
    DevBuf<unsigned char> d_send[2], d_recv[2], d_s_ghost;  // d_recv holds two generations (peer transport)
    // peer transport: neighbours' inboxes mapped through CUDA IPC, messages written in place over NVLink
    bool p2p = false;
    unsigned char* peer_inbox[2] = {nullptr, nullptr};  // [0] = left neighbour's recv[1], [1] = right neighbour's recv[0]
    DevBuf<MsgHeader> d_send_hdr;
    DevBuf<int> d_seq;       // peer transport: device copy of comm_seq, advanced by k_exchange_p2p itself (graph replay)
    unsigned comm_seq = 1;   // sequence number of the next exchange
    unsigned cur_gen = 0;    // inbox generation of the exchange in flight / last completed (fixed at pack time)
    DevBuf<HaloEntry> d_self_ghost;
    DevBuf<int> d_self_ghost_n, d_g_key, d_g_rank;

    // ---- faithful KD-tree neighbour mode (device/kdtree.cuh), allocated when the mode is first selected
    int neighbor_mode = ECMGPU_NEIGHBORS_EXACT;
    DevBuf<unsigned long long> d_kd_keys[2];
    DevBuf<int> d_kd_vals[2], d_kd_seg_r[2], d_kd_seg_node[2], d_kd_raw, d_kd_raw_cnt, d_kd_cache, d_kd_meta;
    DevBuf<float4> d_kd_tree;
    DevBuf<float2> d_kd_pre_pos, d_kd_pre_vel;
    DevBuf<unsigned char> d_kd_sort_tmp;
    size_t kd_sort_tmp_bytes = 0;
    int kd_cap = 0;

    // ---- batched path planning on the device (device/planner.cuh): topology + per-worker scratch, allocated on first use
    DevBuf<float> d_vert_clear;
    DevBuf<int> d_vert_he, d_he_next;
    bool have_topology = false;
    struct PlanBufs {  // per-worker scratch of the device planner (device/planner.cuh: PlanScratch)
        DevBuf<PlanNode> node;
        DevBuf<int> heap, touched, vpath, epath;
        DevBuf<float4> portals;
        DevBuf<float2> out;
        int workers = 0, cap_push = 0, cap_path = 0, cap_portals = 0, cap_out = 0;
        void free() { node.free(); heap.free(); touched.free(); vpath.free(); epath.free(); portals.free(); out.free(); workers = 0; }
    } pl;
    int pl_last_workers = 0, pl_last_second_pass = 0;
    cudaEvent_t pl_ev[2] = {nullptr, nullptr};
    float pl_last_ms = 0.0f;

    // ---- bookkeeping
    uint64_t ticks = 0, launches = 0;
    bool profiling = false;
    bool profile_in_graph = true;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t marks[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // ---- pipelined host I/O (ecmgpu_update_io): two copy streams, double-buffered device staging
    struct IoPipe {
        bool ready = false;
        bool direct_ok = false;  // ECMGPU_IO_DIRECT=1: owned records straight into the caller's pinned buffer (measured slower, see ecmgpu_update_io_owned)
        int cap = 0;
        cudaStream_t s_in = nullptr, s_out = nullptr;
        // staging, two generations each.  Dense calls: in = [pos 8n | vel 8n], out = [pos 8n | vel 8n | active n];
        // owned calls: in = records, out = [32 B header with the count | records]
        unsigned char *in_buf[2] = {nullptr, nullptr}, *out_buf[2] = {nullptr, nullptr};
        // owned calls: where the count of generation b lands on the host, how many records were copied for it
        static constexpr int kTickets = 8;  // calls whose completion can be waited for individually (host pipelines > 2 deep)
        cudaEvent_t ticket_done[kTickets] = {};
        const int32_t* owned_count[kTickets] = {};
        int owned_copied[kTickets] = {};
        uint64_t owned_tick[kTickets] = {};  // ecmgpu_sim::ticks after the call's tick
        // upper bound of the number of agents this handle owns, as far as the host can know it without a
        // synchronisation: the count confirmed by the last ecmgpu_io_wait plus the migrants every tick
        // enqueued since may have brought in; < 0 = unknown (everything is copied)
        long long owned_confirmed = -1;
        uint64_t owned_confirmed_ticket = 0, owned_confirmed_tick = 0;
        cudaEvent_t in_done[2] = {nullptr, nullptr}, in_consumed[2] = {nullptr, nullptr}, tick_done[2] = {nullptr, nullptr}, out_done[2] = {nullptr, nullptr};
        uint64_t calls = 0;
    } io;
    bool ev_valid = false;
    int max_ring = 8;
    // compact walk for strips (device/strips.cuh WalkView): pack / cell count / scatter walk the owned share instead of every
    // slot (13 % off the per-rank tick at 8 x 125 k agents, profiles/r02a_strips8_compact*_launches.csv); env ECMGPU_COMPACT=0 disables
    bool compact = true, walk_dirty = true;
    DevBuf<int> d_walk, d_walk_n;
    DevBuf<unsigned char> d_in_walk;
    // ---- the tick as a CUDA graph (one launch instead of ~15 kernel / memset / NCCL submissions)
    bool use_graph = true;         // env ECMGPU_GRAPH=0 disables
    uint64_t config_epoch = 1;     // bumped whenever something the captured tick depends on changes
    // one graph per inbox generation: the peer transport alternates between two message buffers, everything else
    // in the tick is the same from tick to tick (the exchange sequence number lives in device memory, d_seq)
    uint64_t graph_epoch[2] = {0, 0};
    int graph_n_slots[2] = {-1, -1};
    uint64_t graph_launches[2] = {0, 0};
    cudaGraph_t graph[2] = {nullptr, nullptr};
    cudaGraphExec_t graph_exec[2] = {nullptr, nullptr};
};

namespace {

int fail(ecmgpu_sim* s, int code, const std::string& msg) {
    if (s) s->err = msg;
    else g_create_error = msg;
    return code;
}

#define CUDA_TRY(s, expr)                                                                                   \
    do {                                                                                                    \
        cudaError_t _e = (expr);                                                                            \
        if (_e != cudaSuccess)                                                                              \
            return fail((s), ECMGPU_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));          \
    } while (0)

inline int div_up(int a, int b) { return (a + b - 1) / b; }
// What the per-slot kernels of a tick cover: n_slots rounded up to 4096 (slots beyond n_slots are inactive and cost a
// flag read), so that a host that spawns agents one by one (the drop-in's spawn areas, every tick) does not change the
// launch geometry - and with it re-capture and re-instantiate the tick's CUDA graph - with every new slot.
inline int launch_slots(const ecmgpu_sim* s);
// SMs of a B200: fixed-size grids are multiples of it.  (The host-side test build of this file shrinks it so that its
// thread emulator does not have to start hundreds of thousands of idle threads per launch.)
#ifndef ECM_SM_COUNT
#define ECM_SM_COUNT 148
#endif
constexpr int kSMs = ECM_SM_COUNT;

inline int launch_slots(const ecmgpu_sim* s) { return std::min(s->prm.max_agents, div_up(s->n_slots, 4096) * 4096); }

// ---- host geometry for the static bins --------------------------------------------------------
struct Rect { double x0, y0, x1, y1; };

double point_rect_dist(double px, double py, const Rect& r) {
    double dx = std::max(std::max(r.x0 - px, 0.0), px - r.x1);
    double dy = std::max(std::max(r.y0 - py, 0.0), py - r.y1);
    return std::sqrt(dx * dx + dy * dy);
}
double point_seg_dist(double px, double py, double ax, double ay, double bx, double by) {
    double vx = bx - ax, vy = by - ay, wx = px - ax, wy = py - ay;
    double l2 = vx * vx + vy * vy;
    double t = l2 > 0 ? std::min(1.0, std::max(0.0, (wx * vx + wy * vy) / l2)) : 0.0;
    double cx = ax + t * vx - px, cy = ay + t * vy - py;
    return std::sqrt(cx * cx + cy * cy);
}
bool seg_hits_rect(double ax, double ay, double bx, double by, const Rect& r) {  // Liang-Barsky
    double t0 = 0, t1 = 1, dx = bx - ax, dy = by - ay;
    const double p[4] = {-dx, dx, -dy, dy}, q[4] = {ax - r.x0, r.x1 - ax, ay - r.y0, r.y1 - ay};
    for (int i = 0; i < 4; i++) {
        if (p[i] == 0) { if (q[i] < 0) return false; }
        else {
            double t = q[i] / p[i];
            if (p[i] < 0) { if (t > t1) return false; t0 = std::max(t0, t); }
            else { if (t < t0) return false; t1 = std::min(t1, t); }
        }
    }
    return true;
}
double seg_rect_dist(double ax, double ay, double bx, double by, const Rect& r) {
    if (seg_hits_rect(ax, ay, bx, by, r)) return 0.0;
    double d = std::min(point_rect_dist(ax, ay, r), point_rect_dist(bx, by, r));
    d = std::min(d, point_seg_dist(r.x0, r.y0, ax, ay, bx, by));
    d = std::min(d, point_seg_dist(r.x1, r.y0, ax, ay, bx, by));
    d = std::min(d, point_seg_dist(r.x0, r.y1, ax, ay, bx, by));
    d = std::min(d, point_seg_dist(r.x1, r.y1, ax, ay, bx, by));
    return d;
}

// Builds the two CSR lists of BinView on the host and uploads them.
int build_bins(ecmgpu_sim* s) {
    const double W = (double)s->bbox[2] - s->bbox[0], H = (double)s->bbox[3] - s->bbox[1];
    if (!(W > 0) || !(H > 0)) return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_set_ecm: empty bbox");
    double bin = s->prm.static_bin;
    if (!(bin > 0)) bin = 0.4 * std::sqrt(W * H / std::max(1, s->n_edges));
    // keep the bin count bounded
    while ((W / bin + 3) * (H / bin + 3) > 16.0e6) bin *= 1.5;
    s->static_bin = (float)bin;
    // margin around the walkable area: at least one bin, and at least the obstacle range, so that a point outside the grid
    // is out of reach of every obstacle inside the area (BinView::closed)
    const double range0 = std::max(s->prm.max_obstacle_range > 0 ? (double)s->prm.max_obstacle_range : 0.0, (double)s->tracked_range);
    const double margin = std::ceil(std::max(bin, range0 * 1.0001 + 1e-3 * bin + 2e-3) / bin) * bin;
    s->bins_x0 = (float)(s->bbox[0] - margin);
    s->bins_y0 = (float)(s->bbox[1] - margin);
    s->bins_w = (int)std::ceil((W + 2 * margin) / bin) + 1;
    s->bins_h = (int)std::ceil((H + 2 * margin) / bin) + 1;
    bool closed = true;  // cleared below by a cell that leaves the grid or an obstacle vertex outside the area
    const int nb = s->bins_w * s->bins_h;
    const double x0 = s->bins_x0, y0 = s->bins_y0;
    // The device computes the bin as (int)((p - x0) * inv_bin) in float; pad every footprint by
    // `slack` so float rounding of that expression can never put a point in an unlisted bin.
    const double slack = 1e-3 * bin + 1e-3;
    auto bin_range = [&](double lo, double hi, double origin, int n, int& a, int& b) {
        a = (int)std::floor((lo - slack - origin) / bin);
        b = (int)std::floor((hi + slack - origin) / bin);
        if (a < 0 || b > n - 1 || !(lo == lo) || !(hi == hi)) closed = false;
        a = std::max(a, 0);
        b = std::min(b, n - 1);
    };

    // -- ECM cells: bounding box of the cell polygon (v0, b0, b1, v1)
    std::vector<int> cstart(nb + 1, 0), citems;
    {
        const int nc = 2 * s->n_edges;
        std::vector<int> ax(nc), bx(nc), ay(nc), by(nc);
        for (int c = 0; c < nc; c++) {
            int e = c >> 1, side = c & 1;
            const float* cl = &s->h_edge_cl[8 * e];
            const float* v0 = &s->h_vert_xy[2 * s->h_edge_v[2 * e]];
            const float* v1 = &s->h_vert_xy[2 * s->h_edge_v[2 * e + 1]];
            const float* b0 = cl + 2 * side;
            const float* b1 = cl + 4 + 2 * side;
            double lox = std::min(std::min(v0[0], v1[0]), std::min(b0[0], b1[0]));
            double hix = std::max(std::max(v0[0], v1[0]), std::max(b0[0], b1[0]));
            double loy = std::min(std::min(v0[1], v1[1]), std::min(b0[1], b1[1]));
            double hiy = std::max(std::max(v0[1], v1[1]), std::max(b0[1], b1[1]));
            bin_range(lox, hix, x0, s->bins_w, ax[c], bx[c]);
            bin_range(loy, hiy, y0, s->bins_h, ay[c], by[c]);
            for (int y = ay[c]; y <= by[c]; y++)
                for (int x = ax[c]; x <= bx[c]; x++) cstart[y * s->bins_w + x + 1]++;
        }
        for (int i = 0; i < nb; i++) cstart[i + 1] += cstart[i];
        citems.resize(cstart[nb]);
        std::vector<int> fill(cstart.begin(), cstart.end() - 1);
        for (int c = 0; c < nc; c++)  // ascending c => every list ascending
            for (int y = ay[c]; y <= by[c]; y++)
                for (int x = ax[c]; x <= bx[c]; x++) citems[fill[y * s->bins_w + x]++] = c;
        s->max_cell_list = 0;
        for (int i = 0; i < nb; i++) s->max_cell_list = std::max(s->max_cell_list, cstart[i + 1] - cstart[i]);
    }
    // -- points exactly level with a pass-through cell vertex (BinView::level_hit, host/flat_world.h): the sorted y set, its
    //    hash bitmap, per-row cell lists
    std::vector<float> levels;
    std::vector<unsigned> level_bits;
    std::vector<int> rstart(s->bins_h + 1, 0), ritems;
    {
        const int nc = 2 * s->n_edges;
        for (int c = 0; c < nc; c++) {  // polygon of cell c: v0, boundary.p0, boundary.p1, v1 (ECMCellCollection.cpp:62-80)
            const int e = c >> 1, side = c & 1;
            const float* cl = &s->h_edge_cl[8 * e];
            const float ys[4] = {s->h_vert_xy[2 * s->h_edge_v[2 * e] + 1], cl[2 * side + 1], cl[4 + 2 * side + 1], s->h_vert_xy[2 * s->h_edge_v[2 * e + 1] + 1]};
            ecmb200::pass_through_levels(ys, levels);
        }
        std::sort(levels.begin(), levels.end());
        levels.erase(std::unique(levels.begin(), levels.end()), levels.end());
        int lg = 12;
        while ((1u << lg) < 32u * levels.size() && lg < 24) lg++;
        s->level_shift = 32 - lg;
        level_bits.assign((size_t)1 << (lg - 5), 0u);
        for (float v : levels) {
            unsigned u;
            memcpy(&u, &v, 4);
            const unsigned hsh = (u * 2654435761u) >> s->level_shift;
            level_bits[hsh >> 5] |= 1u << (hsh & 31u);
        }
        s->n_levels = (int)levels.size();
        std::vector<int> ra(nc), rb(nc);
        for (int c = 0; c < nc; c++) {
            const int e = c >> 1, side = c & 1;
            const float* cl = &s->h_edge_cl[8 * e];
            const float ys[4] = {s->h_vert_xy[2 * s->h_edge_v[2 * e] + 1], s->h_vert_xy[2 * s->h_edge_v[2 * e + 1] + 1], cl[2 * side + 1], cl[4 + 2 * side + 1]};
            const double lo = std::min(std::min(ys[0], ys[1]), std::min(ys[2], ys[3])), hi = std::max(std::max(ys[0], ys[1]), std::max(ys[2], ys[3]));
            bin_range(lo, hi, y0, s->bins_h, ra[c], rb[c]);
            for (int y = ra[c]; y <= rb[c]; y++) rstart[y + 1]++;
        }
        for (int i = 0; i < s->bins_h; i++) rstart[i + 1] += rstart[i];
        ritems.resize(rstart[s->bins_h]);
        std::vector<int> fill(rstart.begin(), rstart.end() - 1);
        for (int c = 0; c < nc; c++)  // ascending c => every list ascending
            for (int y = ra[c]; y <= rb[c]; y++) ritems[fill[y]++] = c;
    }
    // -- obstacle segments within `range` of the bin rectangle
    const double range = std::max(s->prm.max_obstacle_range > 0 ? (double)s->prm.max_obstacle_range : 0.0, (double)s->tracked_range);
    std::vector<int> ostart(nb + 1, 0), oitems;
    {
        const int no = s->n_obst;
        const double R = range * 1.0001 + slack;
        std::vector<std::vector<int>> hits(no);
        for (int o = 0; o < no; o++) {
            const double ax_ = s->h_obst_xy[2 * o], ay_ = s->h_obst_xy[2 * o + 1];
            const int nx = s->h_obst_next[o];
            const double bx_ = s->h_obst_xy[2 * nx], by_ = s->h_obst_xy[2 * nx + 1];
            int xa, xb, ya, yb;
            bin_range(std::min(ax_, bx_) - R, std::max(ax_, bx_) + R, x0, s->bins_w, xa, xb);  // leaves the grid => not closed
            bin_range(std::min(ay_, by_) - R, std::max(ay_, by_) + R, y0, s->bins_h, ya, yb);
            for (int y = ya; y <= yb; y++)
                for (int x = xa; x <= xb; x++) {
                    Rect r{x0 + x * bin - slack, y0 + y * bin - slack, x0 + (x + 1) * bin + slack, y0 + (y + 1) * bin + slack};
                    if (seg_rect_dist(ax_, ay_, bx_, by_, r) <= R) {
                        hits[o].push_back(y * s->bins_w + x);
                        ostart[y * s->bins_w + x + 1]++;
                    }
                }
        }
        for (int i = 0; i < nb; i++) ostart[i + 1] += ostart[i];
        oitems.resize(ostart[nb]);
        std::vector<int> fill(ostart.begin(), ostart.end() - 1);
        for (int o = 0; o < no; o++)
            for (int b : hits[o]) oitems[fill[b]++] = o;
        s->max_obst_list = 0;
        for (int i = 0; i < nb; i++) s->max_obst_list = std::max(s->max_obst_list, ostart[i + 1] - ostart[i]);
    }
    CUDA_TRY(s, s->d_bin_cell_start.alloc(nb + 1));
    CUDA_TRY(s, s->d_bin_cell_items.alloc(std::max<size_t>(citems.size(), 1)));
    CUDA_TRY(s, s->d_bin_obst_start.alloc(nb + 1));
    CUDA_TRY(s, s->d_bin_obst_items.alloc(std::max<size_t>(oitems.size(), 1)));
    CUDA_TRY(s, cudaMemcpyAsync(s->d_bin_cell_start.p, cstart.data(), sizeof(int) * (nb + 1), cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(s, cudaMemcpyAsync(s->d_bin_cell_items.p, citems.data(), sizeof(int) * citems.size(), cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(s, cudaMemcpyAsync(s->d_bin_obst_start.p, ostart.data(), sizeof(int) * (nb + 1), cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(s, cudaMemcpyAsync(s->d_bin_obst_items.p, oitems.data(), sizeof(int) * oitems.size(), cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(s, s->d_level_y.alloc(std::max<size_t>(levels.size(), 1)));
    CUDA_TRY(s, s->d_level_bits.alloc(level_bits.size()));
    CUDA_TRY(s, s->d_row_start.alloc(rstart.size()));
    CUDA_TRY(s, s->d_row_items.alloc(std::max<size_t>(ritems.size(), 1)));
    CUDA_TRY(s, cudaMemcpyAsync(s->d_level_y.p, levels.data(), sizeof(float) * levels.size(), cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(s, cudaMemcpyAsync(s->d_level_bits.p, level_bits.data(), sizeof(unsigned) * level_bits.size(), cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(s, cudaMemcpyAsync(s->d_row_start.p, rstart.data(), sizeof(int) * rstart.size(), cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(s, cudaMemcpyAsync(s->d_row_items.p, ritems.data(), sizeof(int) * ritems.size(), cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(s, cudaStreamSynchronize(s->stream));  // the host vectors die here
    s->built_range = (float)range;
    s->bins_closed = closed;
    s->bins_dirty = false;
    s->config_epoch++;
    return ECMGPU_OK;
}

// Chooses the neighbour-grid cell from the crowd's local density and allocates the grid.
int build_grid(ecmgpu_sim* s) {
    double cell = s->prm.neighbor_cell;
    const double W = (double)s->bbox[2] - s->bbox[0], H = (double)s->bbox[3] - s->bbox[1];
    if (!(cell > 0) && s->compact && s->strips_on && s->cell > 0) cell = s->cell;  // keep what the whole crowd chose before it was cut into strips
    if (!(cell > 0)) {
        // local density = agents per occupied 4x4 patch
        std::vector<float2> pos(s->n_slots);
        std::vector<unsigned char> act(s->n_slots);
        CUDA_TRY(s, cudaMemcpyAsync(pos.data(), s->d_pos.p, sizeof(float2) * s->n_slots, cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(s, cudaMemcpyAsync(act.data(), s->d_active.p, s->n_slots, cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(s, cudaStreamSynchronize(s->stream));
        const double patch = 4.0;
        const int pw = std::max(1, (int)std::ceil(W / patch)), ph = std::max(1, (int)std::ceil(H / patch));
        std::vector<unsigned char> occ((size_t)pw * ph, 0);
        size_t n = 0, nocc = 0;
        for (int i = 0; i < s->n_slots; i++) {
            if (!act[i]) continue;
            n++;
            int x = std::min(pw - 1, std::max(0, (int)((pos[i].x - s->bbox[0]) / patch)));
            int y = std::min(ph - 1, std::max(0, (int)((pos[i].y - s->bbox[1]) / patch)));
            if (!occ[(size_t)y * pw + x]) { occ[(size_t)y * pw + x] = 1; nocc++; }
        }
        double rho = n > 0 ? (double)n / ((double)nocc * patch * patch) : 1.0;
        cell = std::min(64.0, std::max(0.5, 1.7 / std::sqrt(rho)));
    }
    while ((W / cell + 3) * (H / cell + 3) > 32.0e6) cell *= 1.5;
    s->cell = (float)cell;
    // Compact strips: the rank's grid covers its strip and halo only, so clearing and scanning the cell table costs the
    // rank's share too.  Exactness does not depend on where the grid lies: agents beyond it are clamped into the border
    // cells and the search treats block sides on the grid border as unbounded (device/knn.cuh).
    double x_lo = s->bbox[0], x_hi = s->bbox[2];
    if (s->compact && s->strips_on) {
        if (s->rank > 0) x_lo = std::max(x_lo, (double)s->strip_lo - s->halo);
        if (s->rank < s->n_ranks - 1) x_hi = std::min(x_hi, (double)s->strip_hi + s->halo);
        if (!(x_hi > x_lo)) { x_lo = s->bbox[0]; x_hi = s->bbox[2]; }
    }
    s->gx0 = (float)(x_lo - cell);
    s->gy0 = (float)(s->bbox[1] - cell);
    s->gw = (int)std::ceil((x_hi - x_lo + 2 * cell) / cell) + 1;
    s->gh = (int)std::ceil((H + 2 * cell) / cell) + 1;
    const int ncells = s->gw * s->gh;
    s->ncells_padded = div_up(ncells + 1, kScanTile) * kScanTile;
    CUDA_TRY(s, s->d_cell_count.alloc(s->ncells_padded));
    CUDA_TRY(s, s->d_scan_state.alloc(s->ncells_padded / kScanTile));
    CUDA_TRY(s, cudaMemsetAsync(s->d_scan_state.p, 0, sizeof(unsigned long long) * s->d_scan_state.n, s->stream));
    if (!s->d_scan_ctl.p) {
        CUDA_TRY(s, s->d_scan_ctl.alloc(4));
        const unsigned ctl0[4] = {1u, 0u, 0u, 0u};  // epoch 1: zeroed states read as "an earlier launch"
        CUDA_TRY(s, cudaMemcpyAsync(s->d_scan_ctl.p, ctl0, sizeof(ctl0), cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(s, cudaStreamSynchronize(s->stream));
    }
    s->grid_dirty = false;
    s->grid_n_slots = std::max(s->n_slots, 1);
    s->config_epoch++;
    return ECMGPU_OK;
}

TickView make_view(ecmgpu_sim* s) {
    TickView t;
    t.ecm.n_vertices = s->n_vertices;
    t.ecm.n_edges = s->n_edges;
    t.ecm.vert_xy = s->d_vert_xy.p;
    t.ecm.edge_v = s->d_edge_v.p;
    t.ecm.edge_cl = s->d_edge_cl.p;
    t.obst.n = s->n_obst;
    t.obst.xy = s->d_obst_xy.p;
    t.obst.next = s->d_obst_next.p;
    t.obst.prev = s->d_obst_prev.p;
    t.obst.convex = s->d_obst_convex.p;
    t.obst.dir = s->d_obst_dir.p;
    t.bins.x0 = s->bins_x0;
    t.bins.y0 = s->bins_y0;
    t.bins.inv_bin = 1.0f / s->static_bin;
    t.bins.w = s->bins_w;
    t.bins.h = s->bins_h;
    t.bins.cell_start = s->d_bin_cell_start.p;
    t.bins.cell_items = s->d_bin_cell_items.p;
    t.bins.obst_start = s->d_bin_obst_start.p;
    t.bins.obst_items = s->d_bin_obst_items.p;
    t.bins.closed = s->bins_closed ? 1 : 0;
    t.bins.level_y = s->d_level_y.p;
    t.bins.level_bits = s->d_level_bits.p;
    t.bins.n_levels = s->n_levels;
    t.bins.level_shift = s->level_shift;
    t.bins.row_start = s->d_row_start.p;
    t.bins.row_items = s->d_row_items.p;
    t.grid.x0 = s->gx0;
    t.grid.y0 = s->gy0;
    t.grid.cell = s->cell;
    t.grid.inv_cell = 1.0f / s->cell;
    t.grid.w = s->gw;
    t.grid.h = s->gh;
    t.grid.n_sorted = 0;
    t.grid.cell_start = s->d_cell_count.p;
    t.grid.s_pos = s->d_s_pos.p;
    t.grid.s_vel = s->d_s_vel.p;
    t.grid.s_rad = s->d_s_rad.p;
    t.grid.s_slot = s->d_s_slot.p;
    t.grid.ext_of = s->perm_identity ? nullptr : s->d_ext_of.p;
    t.ag.pos = s->d_pos.p;
    t.ag.vel = s->d_vel.p;
    t.ag.prefvel = s->d_prefvel.p;
    t.ag.attraction = s->d_attraction.p;
    t.ag.force = s->d_force.p;
    t.ag.radius = s->d_radius.p;
    t.ag.speed = s->d_speed.p;
    t.ag.active = s->d_active.p;
    t.ag.replan_pending = s->d_replan_pending.p;
    t.ag.status = s->d_status.p;
    t.ag.cell = s->d_cell.p;
    t.ag.nbr = s->d_nbr.p;
    t.ag.nbr_cnt = s->d_nbr_cnt.p;
    t.ag.path_hdr = s->d_path_hdr.p;
    t.ag.path_pool = s->d_path_pool.p;
    t.ag.path_bbox = s->d_path_bbox.p;
    t.sc.key = s->d_key.p;
    t.sc.rank = s->d_rank.p;
    t.sc.cell_count = s->d_cell_count.p;
    t.sc.s_pos = s->d_s_pos.p;
    t.sc.s_vel = s->d_s_vel.p;
    t.sc.s_rad = s->d_s_rad.p;
    t.sc.s_spd = s->d_s_spd.p;
    t.sc.s_slot = s->d_s_slot.p;
    t.sc.s_pref = s->d_s_pref.p;
    t.sc.s_alive = s->d_s_alive.p;
    t.sc.s_ghost = s->d_s_ghost.p;
    t.sc.fb_list = s->d_fb_list.p;
    t.sc.ev_replan = s->d_ev_replan.p;
    t.sc.ev_destroyed = s->d_ev_destroyed.p;
    t.sc.ev_cap = (int)s->d_ev_destroyed.n;
    t.sc.counters = s->d_counters.p;
    t.n_sorted_ptr = s->d_cell_count.p + (size_t)s->gw * s->gh;
    t.step = s->prm.step;
    t.max_ring = s->max_ring;
    t.record_neighbors = s->prm.record_neighbors;
    t.strips = s->strips_on ? 1 : 0;
    const float inf = std::numeric_limits<float>::infinity();
    t.cover_lo = s->strips_on && s->rank > 0 ? s->strip_lo - s->halo : -inf;
    t.cover_hi = s->strips_on && s->rank < s->n_ranks - 1 ? s->strip_hi + s->halo : inf;
    t.lp3d.cap = s->lp3d_cap;
    t.lp3d.count = s->d_counters.p + C_LP3D_N;
    t.lp3d.hdr = s->d_lp3d_hdr.p;
    t.lp3d.out = s->d_lp3d_out.p;
    t.lp3d.cs = s->d_lp3d_cs.p;
    return t;
}

int ensure_ready(ecmgpu_sim* s) {
    if (!s->have_ecm) return fail(s, ECMGPU_ERR_INVALID, "no ECM: call ecmgpu_set_ecm first");
    const float want = std::max(s->prm.max_obstacle_range > 0 ? s->prm.max_obstacle_range : 0.0f, s->tracked_range);
    if (s->bins_dirty || want > s->built_range) {
        int rc = build_bins(s);
        if (rc) return rc;
    }
    // The automatic cell edge comes from the crowd's density when the grid is built.  A simulator that starts empty and
    // fills through spawn areas would keep the cell chosen for its first handful of agents (exact, but 10-20 x the
    // candidates per search): choose again whenever the loaded slots have doubled (or halved) since.
    if (!s->grid_dirty && !(s->prm.neighbor_cell > 0) && !s->strips_on && s->grid_n_slots > 0 &&
        (s->n_slots > 2 * s->grid_n_slots || 2 * s->n_slots < s->grid_n_slots))
        s->grid_dirty = true;
    if (s->grid_dirty) {
        int rc = build_grid(s);
        if (rc) return rc;
    }
    return ECMGPU_OK;
}

// ---- NCCL, loaded lazily so that single-GPU users need no NCCL at all ----------------------------
struct UniqueId { char internal[128]; };  // ncclUniqueId
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, UniqueId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

bool load_nccl(std::string& err) {
    if (g_nccl.lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) { err = std::string("cannot load libnccl: ") + dlerror(); return false; }
    auto sym = [&](const char* n) { void* p = dlsym(h, n); if (!p) err = std::string("missing NCCL symbol ") + n; return p; };
    g_nccl.GetUniqueId = (int (*)(void*))sym("ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void**, int, UniqueId, int))sym("ncclCommInitRank");
    g_nccl.CommDestroy = (int (*)(void*))sym("ncclCommDestroy");
    g_nccl.Send = (int (*)(const void*, size_t, int, int, void*, cudaStream_t))sym("ncclSend");
    g_nccl.Recv = (int (*)(void*, size_t, int, int, void*, cudaStream_t))sym("ncclRecv");
    g_nccl.GroupStart = (int (*)())sym("ncclGroupStart");
    g_nccl.GroupEnd = (int (*)())sym("ncclGroupEnd");
    g_nccl.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    if (!err.empty()) return false;
    g_nccl.lib = h;
    return true;
}

#define NCCL_TRY(s, expr)                                                                                  \
    do {                                                                                                   \
        int _r = (expr);                                                                                   \
        if (_r != 0) return fail((s), ECMGPU_ERR_COMM, std::string(#expr) + ": " + g_nccl.GetErrorString(_r)); \
    } while (0)

void comm_teardown(ecmgpu_sim* s) {
    if (s->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(s->nccl_comm);
    s->nccl_comm = nullptr;
    if (s->ev_packed) cudaEventDestroy(s->ev_packed);
    if (s->ev_pulled) cudaEventDestroy(s->ev_pulled);
    s->ev_packed = s->ev_pulled = nullptr;
    for (int d = 0; d < 2; d++) {
        if (s->peer_inbox[d]) cudaIpcCloseMemHandle(s->peer_inbox[d]);
        s->peer_inbox[d] = nullptr;
        s->d_send[d].free();
        s->d_recv[d].free();
    }
    s->d_send_hdr.free();
    s->d_seq.free();
    s->p2p = false;
    s->d_s_ghost.free(); s->d_self_ghost.free(); s->d_self_ghost_n.free(); s->d_g_key.free(); s->d_g_rank.free();
}

StripView make_strip_view(ecmgpu_sim* s) {
    StripView v;
    memset(&v, 0, sizeof(v));
    v.enabled = s->strips_on ? 1 : 0;
    v.rank = s->rank;
    v.n_ranks = s->n_ranks;
    v.lo = s->strip_lo;
    v.hi = s->strip_hi;
    v.halo = s->halo;
    v.cap_halo = s->cap_halo;
    v.cap_migr = s->cap_migr;
    v.cap_self = s->cap_self;
    const size_t msg = strip_msg_bytes(s->cap_halo, s->cap_migr);
    const size_t gen = s->p2p ? (size_t)s->cur_gen * msg : 0;
    for (int d = 0; d < 2; d++) {
        v.recv[d] = s->d_recv[d].p ? s->d_recv[d].p + gen : nullptr;
        if (s->p2p) {
            v.send[d] = s->peer_inbox[d] ? s->peer_inbox[d] + gen : s->d_send[d].p;
            v.send_hdr[d] = s->d_send_hdr.p + d;
        } else {
            v.send[d] = s->d_send[d].p;
            v.send_hdr[d] = (MsgHeader*)s->d_send[d].p;
        }
    }
    v.self_ghost = s->d_self_ghost.p;
    v.self_ghost_n = s->d_self_ghost_n.p;
    v.g_key = s->d_g_key.p;
    v.g_rank = s->d_g_rank.p;
    const bool walk = s->compact && s->strips_on && s->d_walk.p;
    v.walk.list = walk ? s->d_walk.p : nullptr;
    v.walk.n = s->d_walk_n.p;
    v.walk.in_list = s->d_in_walk.p;
    return v;
}

// Compact walk: (re)built from the active flags whenever the host changed who owns what (loads, spawns, writes of the
// active flags, new strip borders); between rebuilds the device keeps it current (adopted migrants are appended).
int ensure_walk(ecmgpu_sim* s) {
    if (!s->compact || !s->strips_on) return ECMGPU_OK;
    if (!s->d_walk.p) {
        const size_t n = (size_t)s->prm.max_agents;
        CUDA_TRY(s, s->d_walk.alloc(n)); CUDA_TRY(s, s->d_walk_n.alloc(1)); CUDA_TRY(s, s->d_in_walk.alloc(n));
        s->walk_dirty = true;
    }
    if (!s->walk_dirty) return ECMGPU_OK;
    CUDA_TRY(s, cudaMemsetAsync(s->d_walk_n.p, 0, sizeof(int), s->stream));
    CUDA_TRY(s, cudaMemsetAsync(s->d_in_walk.p, 0, (size_t)s->prm.max_agents, s->stream));
    if (s->n_slots > 0) {
        WalkView w{s->d_walk.p, s->d_walk_n.p, s->d_in_walk.p};
        k_walk_rebuild<<<div_up(s->n_slots, kPackBlock), kPackBlock, 0, s->stream>>>(s->n_slots, s->d_active.p, w);
        s->launches++;
        CUDA_TRY(s, cudaGetLastError());
    }
    s->walk_dirty = false;
    return ECMGPU_OK;
}

// phase 0: pack halo / migrant / self-ghost lists from the current state
int enqueue_pack(ecmgpu_sim* s, const TickView& t) {
    int rc = ensure_walk(s);  // a no-op inside a graph capture: ecmgpu_update has done it before the capture began
    if (rc) return rc;
    s->cur_gen = s->comm_seq & 1u;
    StripView sv = make_strip_view(s);
    if (s->local_transport)  // neighbours must have pulled last tick's messages before we overwrite them
        for (int d = 0; d < 2; d++)
            if (s->peer[d]) CUDA_TRY(s, cudaStreamWaitEvent(s->stream, s->peer[d]->ev_pulled, 0));
    if (s->p2p) CUDA_TRY(s, cudaMemsetAsync(s->d_send_hdr.p, 0, 2 * sizeof(MsgHeader), s->stream));
    else for (int d = 0; d < 2; d++) CUDA_TRY(s, cudaMemsetAsync(s->d_send[d].p, 0, sizeof(MsgHeader), s->stream));
    CUDA_TRY(s, cudaMemsetAsync(s->d_self_ghost_n.p, 0, sizeof(int), s->stream));
    if (sv.walk.list) k_pack_walk<<<kSMs * 2, kPackBlock, 0, s->stream>>>(t.ag, sv, s->d_counters.p);
    else k_pack<<<div_up(launch_slots(s), kPackBlock), kPackBlock, 0, s->stream>>>(launch_slots(s), t.ag, sv, s->d_counters.p);
    s->launches++;
    if (s->local_transport) CUDA_TRY(s, cudaEventRecord(s->ev_packed, s->stream));
    CUDA_TRY(s, cudaGetLastError());
    return ECMGPU_OK;
}

// phase 1: one fixed-size message per direction (NCCL send/recv, or peer copies inside one process),
// then received migrants become owned agents
int enqueue_exchange(ecmgpu_sim* s, const TickView& t) {
    StripView sv = make_strip_view(s);
    const size_t msg = strip_msg_bytes(s->cap_halo, s->cap_migr);
    if (s->local_transport) {
        for (int d = 0; d < 2; d++) {
            ecmgpu_sim* p = s->peer[d];
            if (!p) continue;
            CUDA_TRY(s, cudaStreamWaitEvent(s->stream, p->ev_packed, 0));
            // my left neighbour's RIGHT message is my left inbox, and vice versa
            CUDA_TRY(s, cudaMemcpyPeerAsync(s->d_recv[d].p, s->prm.device, p->d_send[1 - d].p, p->prm.device, msg, s->stream));
        }
        CUDA_TRY(s, cudaEventRecord(s->ev_pulled, s->stream));
    } else if (s->p2p) {
        k_exchange_p2p<<<1, 256, 0, s->stream>>>(sv, t.ag, s->d_seq.p);  // publish, await, adopt the migrants
        s->launches++;
    } else if (s->n_ranks > 1) {
        NCCL_TRY(s, g_nccl.GroupStart());
        if (s->rank > 0) {
            NCCL_TRY(s, g_nccl.Send(s->d_send[0].p, msg, /*ncclInt8*/ 0, s->rank - 1, s->nccl_comm, s->stream));
            NCCL_TRY(s, g_nccl.Recv(s->d_recv[0].p, msg, 0, s->rank - 1, s->nccl_comm, s->stream));
        }
        if (s->rank < s->n_ranks - 1) {
            NCCL_TRY(s, g_nccl.Send(s->d_send[1].p, msg, 0, s->rank + 1, s->nccl_comm, s->stream));
            NCCL_TRY(s, g_nccl.Recv(s->d_recv[1].p, msg, 0, s->rank + 1, s->nccl_comm, s->stream));
        }
        NCCL_TRY(s, g_nccl.GroupEnd());
    }
    if (!s->p2p) {
        k_unpack_migrants<<<div_up(std::max(s->cap_migr, 1), 256), 256, 0, s->stream>>>(t.ag, sv);
        s->launches++;
    }
    s->comm_seq++;  // the next exchange uses the other inbox generation
    CUDA_TRY(s, cudaGetLastError());
    return ECMGPU_OK;
}

// count + scan + scatter: the per-tick neighbour structure (owned agents, plus ghosts with strips)
int enqueue_grid_build(ecmgpu_sim* s, const TickView& t) {
    GridParams gp{s->gx0, s->gy0, s->cell, 1.0f / s->cell, s->gw, s->gh};
    CUDA_TRY(s, cudaMemsetAsync(s->d_cell_count.p, 0, sizeof(int) * s->ncells_padded, s->stream));
    static_assert(C_LP3D_N == C_FALLBACK_N + 1, "the two per-tick counters are cleared by one memset");
    CUDA_TRY(s, cudaMemsetAsync(s->d_counters.p + C_FALLBACK_N, 0, 2 * sizeof(unsigned long long), s->stream));
    const int nls = launch_slots(s), nb = div_up(nls, 256);
    StripView sv = make_strip_view(s);
    const int ng = 2 * s->cap_halo + s->cap_self;
    // compact strips: the list-walking kernels take the ghosts along (one launch each instead of two)
    if (sv.walk.list) k_bin_count_walk<<<kSMs * 8, 256, 0, s->stream>>>(sv, s->d_active.p, s->d_pos.p, gp, s->d_cell_count.p, s->d_key.p, s->d_rank.p, s->d_status.p, s->d_counters.p);
    else {
        k_bin_count<<<nb, 256, 0, s->stream>>>(nls, s->d_active.p, s->d_pos.p, gp, s->d_cell_count.p, s->d_key.p, s->d_rank.p, s->d_status.p, s->d_counters.p);
        if (s->strips_on) {
            k_ghost_count<<<div_up(ng, 256), 256, 0, s->stream>>>(sv, gp, s->d_cell_count.p);
            s->launches++;
        }
    }
    const int tiles = s->ncells_padded / kScanTile;
    k_scan_onepass<<<tiles, kScanBlock, 0, s->stream>>>((int4*)s->d_cell_count.p, tiles, s->d_scan_state.p, s->d_scan_ctl.p);
    if (sv.walk.list) k_scatter_walk<<<kSMs * 8, 256, 0, s->stream>>>(sv, s->d_key.p, s->d_rank.p, s->d_cell_count.p, t.ag, t.sc);
    else {
        k_scatter<<<nb, 256, 0, s->stream>>>(nls, s->d_key.p, s->d_rank.p, s->d_cell_count.p, t.ag, t.sc);
        if (s->strips_on) {
            k_ghost_scatter<<<div_up(ng, 256), 256, 0, s->stream>>>(sv, s->d_cell_count.p, t.ag, t.sc);
            s->launches++;
        }
    }
    s->launches += 3;
    CUDA_TRY(s, cudaGetLastError());
    return ECMGPU_OK;
}

// ---- batched path planning (device/planner.cuh) ----------------------------------------------------
void plan_free(ecmgpu_sim* s) { s->pl.free(); }

PlanScratch make_plan_scratch(ecmgpu_sim::PlanBufs& b) {
    PlanScratch sc;
    sc.n_workers = b.workers;
    sc.cap_push = b.cap_push; sc.cap_path = b.cap_path; sc.cap_portals = b.cap_portals; sc.cap_out = b.cap_out;
    sc.node = b.node.p; sc.heap = b.heap.p; sc.touched = b.touched.p; sc.vpath = b.vpath.p; sc.epath = b.epath.p;
    sc.portals = b.portals.p; sc.out = b.out.p;
    return sc;
}

// Per-worker scratch for `want` concurrent queries, within a memory budget: half of the free device memory, at most
// 96 GB (first pass; env ECMGPU_PLAN_MB overrides; the 4 M-agent map's graph needs 1.5 MB per query in flight) / a
// quarter, at most 8 GB (second pass).  ecmgpu_plan_paths keeps up to 2 GB of it between calls.
//   full = true:  2E + 4 pushes (a query pushes at most once per directed edge, plus the two start vertices) and the
//                 capacities include/ecm_b200.h documents (2048 graph vertices, 8192 portals, 1024 points).
//   full = false: the first pass: 2048 portals, and the same room for pushes as long as 32 k such workers fit the
//                 budget - the kernel is bound by the sector rate of DRAM and gains nothing beyond ~100 k queries in
//                 flight, but loses below ~50 k (profiles/r03b_planner_probe.jsonl).  On a graph too large for that
//                 the push capacity shrinks (>= 8192) and the queries that fill it are planned again in the second
//                 pass (1.4 % of the C3 crowd's routes push more than 8192 times, 14 % more than 4096:
//                 profiles/r03c_planner_sweep.jsonl).  Env ECMGPU_PLAN_PUSH sets the first-pass capacity.
int plan_alloc(ecmgpu_sim* s, ecmgpu_sim::PlanBufs& b, int want, bool full) {
    const size_t nV = (size_t)s->n_vertices, nE = (size_t)s->n_edges;
    const int cap_path = (int)std::min<size_t>(nV + 2, 2048);  // vertices of one A* path
    const int cap_portals = full ? 8192 : 2048, cap_out = 1024;
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(s, cudaMemGetInfo(&free_b, &total_b));
    size_t budget = full ? std::min<size_t>(free_b / 4, (size_t)8 << 30) : std::min<size_t>(free_b / 2, (size_t)96 << 30);
    if (!full) if (const char* e = getenv("ECMGPU_PLAN_MB")) budget = (size_t)std::max(16, atoi(e)) << 20;
    const size_t fixed = nV * sizeof(PlanNode) + (size_t)cap_path * 8 + (size_t)cap_portals * 16 + (size_t)cap_out * 8;
    size_t push = 2 * nE + 4;
    if (!full) {
        const size_t per_worker_min = budget / 32768;
        if (fixed + push * 8 > per_worker_min) push = std::min(push, std::max<size_t>(8192, per_worker_min > fixed ? (per_worker_min - fixed) / 8 : 0));
        if (const char* e = getenv("ECMGPU_PLAN_PUSH")) push = std::min<size_t>(2 * nE + 4, (size_t)std::max(64, atoi(e)));
    }
    const int cap_push = (int)std::min<size_t>(push, (size_t)INT32_MAX);
    const size_t per_worker = fixed + (size_t)cap_push * 8;
    int workers = (int)std::min<size_t>({(size_t)want, (size_t)kSMs * kPlanThreadsPerSM, std::max<size_t>(budget / per_worker, 32)});
    workers = div_up(workers, 32) * 32;
    if (b.workers >= workers && b.cap_push == cap_push && b.cap_portals == cap_portals) return ECMGPU_OK;
    b.free();
    const size_t w = (size_t)workers;
    CUDA_TRY(s, b.node.alloc(w * nV));
    CUDA_TRY(s, b.heap.alloc(w * cap_push)); CUDA_TRY(s, b.touched.alloc(w * cap_push));
    CUDA_TRY(s, b.vpath.alloc(w * cap_path)); CUDA_TRY(s, b.epath.alloc(w * cap_path));
    CUDA_TRY(s, b.portals.alloc(w * cap_portals)); CUDA_TRY(s, b.out.alloc(w * cap_out));
    b.workers = workers;
    b.cap_push = cap_push; b.cap_path = cap_path; b.cap_portals = cap_portals; b.cap_out = cap_out;
    k_plan_init<<<kSMs * 8, 256, 0, s->stream>>>(make_plan_scratch(b), (int)nV);
    s->launches++;
    CUDA_TRY(s, cudaGetLastError());
    return ECMGPU_OK;
}

int ensure_bins(ecmgpu_sim* s) {
    if (!s->have_ecm) return fail(s, ECMGPU_ERR_INVALID, "no ECM: call ecmgpu_set_ecm first");
    const float want = std::max(s->prm.max_obstacle_range > 0 ? s->prm.max_obstacle_range : 0.0f, s->tracked_range);
    if (s->bins_dirty || want > s->built_range) return build_bins(s);
    return ECMGPU_OK;
}

// ---- faithful KD-tree neighbour mode (device/kdtree.cuh) -----------------------------------------
// Levels of a median-split tree over n agents: the smallest L with 2^L - 1 >= n.
int kd_levels(int n) {
    int L = 0;
    while (((1ll << L) - 1) < (long long)n) L++;
    return L;
}

int kd_alloc(ecmgpu_sim* s) {
    if (s->d_kd_raw.p) return ECMGPU_OK;
    const size_t n = (size_t)s->prm.max_agents;
    for (int b = 0; b < 2; b++) {
        CUDA_TRY(s, s->d_kd_keys[b].alloc(n)); CUDA_TRY(s, s->d_kd_vals[b].alloc(n));
        CUDA_TRY(s, s->d_kd_seg_r[b].alloc(n + 1)); CUDA_TRY(s, s->d_kd_seg_node[b].alloc(n + 1));
    }
    s->kd_cap = (int)((1ll << kd_levels((int)n)) - 1);
    CUDA_TRY(s, s->d_kd_tree.alloc((size_t)s->kd_cap));
    CUDA_TRY(s, s->d_kd_raw.alloc(5 * n)); CUDA_TRY(s, s->d_kd_raw_cnt.alloc(n));
    CUDA_TRY(s, s->d_kd_cache.alloc(10));  // [0,5): carried from tick to tick; [5,10): zeros for ecmgpu_find_neighbors
    CUDA_TRY(s, s->d_kd_meta.alloc(4));
    CUDA_TRY(s, s->d_kd_pre_pos.alloc(n)); CUDA_TRY(s, s->d_kd_pre_vel.alloc(n));
    cub::DoubleBuffer<unsigned long long> dk(s->d_kd_keys[0].p, s->d_kd_keys[1].p);
    cub::DoubleBuffer<int> dv(s->d_kd_vals[0].p, s->d_kd_vals[1].p);
    size_t bytes = 0;
    CUDA_TRY(s, cub::DeviceRadixSort::SortPairs(nullptr, bytes, dk, dv, (int)n, 0, 64, s->stream));
    s->kd_sort_tmp_bytes = bytes;
    CUDA_TRY(s, s->d_kd_sort_tmp.alloc(std::max<size_t>(bytes, 16)));
    CUDA_TRY(s, cudaMemsetAsync(s->d_kd_meta.p, 0, sizeof(int) * 4, s->stream));
    return ECMGPU_OK;
}

void kd_free(ecmgpu_sim* s) {
    for (int b = 0; b < 2; b++) { s->d_kd_keys[b].free(); s->d_kd_vals[b].free(); s->d_kd_seg_r[b].free(); s->d_kd_seg_node[b].free(); }
    s->d_kd_tree.free(); s->d_kd_raw.free(); s->d_kd_raw_cnt.free(); s->d_kd_cache.free(); s->d_kd_meta.free();
    s->d_kd_pre_pos.free(); s->d_kd_pre_vel.free(); s->d_kd_sort_tmp.free();
}

// KDTree::Construct (KDTree.cpp:22-57) on the pre-tick positions: one segmented sort + one split kernel per level.
// Needs the snapshot's slot list (enqueue_grid_build) for the agents active at the start of the tick.
int enqueue_kd_build(ecmgpu_sim* s, const TickView& t) {
    const int n = s->n_slots;
    KdBuild b;
    b.n_slots = n;
    b.n_active_ptr = t.n_sorted_ptr;
    b.cell_key = s->d_key.p;
    b.pos = s->d_pos.p;
    b.tree = s->d_kd_tree.p;
    b.cap = s->kd_cap;
    b.meta = s->d_kd_meta.p;
    b.ties = s->d_counters.p + C_TOTAL_KD_TIES;
    b.small_ties = s->d_counters.p + C_TOTAL_KD_SMALL_TIES;
    const int levels = kd_levels(n);
    const size_t used_nodes = (size_t)((1ll << levels) - 1);
    CUDA_TRY(s, cudaMemsetAsync(s->d_kd_tree.p, 0xff, sizeof(float4) * used_nodes, s->stream));  // KDTREE_NULL_NODE everywhere (KDTree.cpp:50-51)
    const int nb = div_up(n, 256);
    unsigned long long* k_in = s->d_kd_keys[0].p; unsigned long long* k_alt = s->d_kd_keys[1].p;
    int* v_in = s->d_kd_vals[0].p; int* v_alt = s->d_kd_vals[1].p;
    k_kd_init<<<nb, 256, 0, s->stream>>>(b, k_in, v_in, s->d_kd_seg_r[0].p, s->d_kd_seg_node[0].p);
    s->launches++;
    int seg_bits = 1;  // segment starts are < n, the dead elements of level 0 carry n itself
    while ((1ll << seg_bits) <= (long long)n) seg_bits++;
    for (int d = 0; d < levels; d++) {
        cub::DoubleBuffer<unsigned long long> dk(k_in, k_alt);
        cub::DoubleBuffer<int> dv(v_in, v_alt);
        size_t bytes = s->kd_sort_tmp_bytes;
        CUDA_TRY(s, cub::DeviceRadixSort::SortPairs(s->d_kd_sort_tmp.p, bytes, dk, dv, n, 0, 32 + seg_bits, s->stream));
        k_kd_split<<<nb, 256, 0, s->stream>>>(b, d, dk.Current(), dv.Current(), dk.Alternate(), dv.Alternate(), s->d_kd_seg_r[d & 1].p,
                                              s->d_kd_seg_node[d & 1].p, s->d_kd_seg_r[(d + 1) & 1].p, s->d_kd_seg_node[(d + 1) & 1].p);
        k_in = dk.Alternate(); k_alt = dk.Current();
        v_in = dv.Alternate(); v_alt = dv.Current();
        s->launches++;  // ours; the library's sort passes are not counted
    }
    CUDA_TRY(s, cudaGetLastError());
    return ECMGPU_OK;
}

KdQuery make_kd_query(ecmgpu_sim* s, bool carried) {
    KdQuery q;
    q.tree = s->d_kd_tree.p;
    q.cap = s->kd_cap;
    q.meta = s->d_kd_meta.p;
    q.raw = s->d_kd_raw.p;
    q.raw_cnt = s->d_kd_raw_cnt.p;
    q.cache = s->d_kd_cache.p + (carried ? 0 : 5);
    return q;
}

// The ORCA phase of a tick in KD-tree mode: tree, lists, token resolution, carried list, k_orca_kd.
int enqueue_kd_orca(ecmgpu_sim* s, const TickView& t, int rows) {
    int rc = enqueue_kd_build(s, t);
    if (rc) return rc;
    const size_t n = (size_t)s->n_slots;
    // neighbours are read by slot from the pre-tick state while k_orca_kd integrates in place
    CUDA_TRY(s, cudaMemcpyAsync(s->d_kd_pre_pos.p, s->d_pos.p, sizeof(float2) * n, cudaMemcpyDeviceToDevice, s->stream));
    CUDA_TRY(s, cudaMemcpyAsync(s->d_kd_pre_vel.p, s->d_vel.p, sizeof(float2) * n, cudaMemcpyDeviceToDevice, s->stream));
    const KdQuery q = make_kd_query(s, true);
    k_kd_query<<<div_up(rows, 128), 128, 0, s->stream>>>(t, q, 1);
    k_kd_resolve<<<div_up(s->n_slots, 256), 256, 0, s->stream>>>(s->n_slots, s->d_active.p, q, s->d_nbr.p, s->d_nbr_cnt.p);
    k_kd_cache<<<1, 256, 0, s->stream>>>(s->n_slots, s->d_active.p, s->d_nbr.p, q.cache);
    TickView t2 = t;
    t2.grid.s_pos = s->d_kd_pre_pos.p;
    t2.grid.s_vel = s->d_kd_pre_vel.p;
    t2.grid.s_rad = s->d_radius.p;
    t2.record_neighbors = 0;  // ECMGPU_NEIGHBORS already holds the lists
    k_orca_kd<<<div_up(rows, 256), 256, 0, s->stream>>>(t2);
    s->launches += 4;
    CUDA_TRY(s, cudaGetLastError());
    return ECMGPU_OK;
}


// Appends a polyline to the host mirror of the pool (start aligned to 8 points) with the padded
// bounding boxes of its blocks of 8 segments; compacts the pool when it is full.
void append_path(ecmgpu_sim* s, int slot, const float2* pts, int n) {
    while (s->h_path_pool.size() % kPathBlock) s->h_path_pool.push_back(make_float2(0.0f, 0.0f));
    const int off = (int)s->h_path_pool.size();
    s->h_path_pool.insert(s->h_path_pool.end(), pts, pts + n);
    const int nseg = n - 1, nblk = (nseg + kPathBlock - 1) / kPathBlock;
    s->h_path_bbox.resize((size_t)off / kPathBlock + std::max(nblk, 1), make_float4(0, 0, 0, 0));
    const float pad = 0.05f;  // far above the rounding of the reference's intersection arithmetic
    for (int b = 0; b < nblk; b++) {
        float4 bb = make_float4(3e38f, 3e38f, -3e38f, -3e38f);
        for (int i = b * kPathBlock; i <= std::min((b + 1) * kPathBlock, nseg); i++) {
            bb.x = std::min(bb.x, pts[i].x); bb.y = std::min(bb.y, pts[i].y);
            bb.z = std::max(bb.z, pts[i].x); bb.w = std::max(bb.w, pts[i].y);
        }
        bb.x -= pad; bb.y -= pad; bb.z += pad; bb.w += pad;
        s->h_path_bbox[(size_t)off / kPathBlock + b] = bb;
    }
    s->h_path_hdr[slot] = PathHdr{off, n, pts[n - 1].x, pts[n - 1].y};
}

int upload_path(ecmgpu_sim* s, int slot, const float* xy, int n) {
    if (n < 1) return fail(s, ECMGPU_ERR_INVALID, "path needs at least 1 point");
    const size_t cap = s->d_path_pool.n;
    if (s->h_path_pool.size() + n + kPathBlock > cap) {
        // compact: rebuild the pool from the live headers
        std::vector<float2> old;
        old.swap(s->h_path_pool);
        s->h_path_bbox.clear();
        for (int i = 0; i < s->n_slots; i++) {
            const PathHdr h = s->h_path_hdr[i];
            if (h.len <= 0 || i == slot) { if (i == slot) s->h_path_hdr[i] = PathHdr{0, 0, 0.0f, 0.0f}; continue; }
            std::vector<float2> tmp(old.begin() + h.off, old.begin() + h.off + h.len);
            append_path(s, i, tmp.data(), h.len);
        }
        if (s->h_path_pool.size() + n + kPathBlock > cap) return fail(s, ECMGPU_ERR_CAPACITY, "path pool full (raise ecmgpu_params.path_pool_points)");
        CUDA_TRY(s, cudaMemcpyAsync(s->d_path_pool.p, s->h_path_pool.data(), sizeof(float2) * s->h_path_pool.size(), cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(s, cudaMemcpyAsync(s->d_path_bbox.p, s->h_path_bbox.data(), sizeof(float4) * s->h_path_bbox.size(), cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(s, cudaMemcpyAsync(s->d_path_hdr.p, s->h_path_hdr.data(), sizeof(PathHdr) * s->n_slots, cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(s, cudaStreamSynchronize(s->stream));
        s->pool_uploaded = s->h_path_pool.size();
    }
    std::vector<float2> pts(n);
    for (int j = 0; j < n; j++) pts[j] = make_float2(xy[2 * j], xy[2 * j + 1]);
    append_path(s, slot, pts.data(), n);
    return ECMGPU_OK;
}

// Uploads the part of the pool mirror (points + bboxes) appended since the last upload.
int flush_pool(ecmgpu_sim* s) {
    if (s->h_path_pool.size() > s->pool_uploaded) {
        const size_t a = s->pool_uploaded - s->pool_uploaded % kPathBlock, e = s->h_path_pool.size();
        CUDA_TRY(s, cudaMemcpyAsync(s->d_path_pool.p + a, s->h_path_pool.data() + a, sizeof(float2) * (e - a), cudaMemcpyHostToDevice, s->stream));
        const size_t ba = a / kPathBlock, be = s->h_path_bbox.size();
        if (be > ba) CUDA_TRY(s, cudaMemcpyAsync(s->d_path_bbox.p + ba, s->h_path_bbox.data() + ba, sizeof(float4) * (be - ba), cudaMemcpyHostToDevice, s->stream));
        s->pool_uploaded = e;
    }
    return ECMGPU_OK;
}

size_t elem_size(int which) {
    switch (which) {
        case ECMGPU_POS: case ECMGPU_VEL: case ECMGPU_PREFVEL: case ECMGPU_ATTRACTION: case ECMGPU_FORCE: return 8;
        case ECMGPU_RADIUS: case ECMGPU_SPEED: case ECMGPU_CELL: case ECMGPU_NEIGHBOR_COUNT: case ECMGPU_STATUS: return 4;
        case ECMGPU_ACTIVE: case ECMGPU_REPLAN_PENDING: return 1;
        case ECMGPU_NEIGHBORS: return 20;
        default: return 0;
    }
}
void* dev_array(ecmgpu_sim* s, int which) {
    switch (which) {
        case ECMGPU_POS: return s->d_pos.p;
        case ECMGPU_VEL: return s->d_vel.p;
        case ECMGPU_PREFVEL: return s->d_prefvel.p;
        case ECMGPU_ATTRACTION: return s->d_attraction.p;
        case ECMGPU_FORCE: return s->d_force.p;
        case ECMGPU_RADIUS: return s->d_radius.p;
        case ECMGPU_SPEED: return s->d_speed.p;
        case ECMGPU_ACTIVE: return s->d_active.p;
        case ECMGPU_CELL: return s->d_cell.p;
        case ECMGPU_NEIGHBORS: return s->d_nbr.p;
        case ECMGPU_NEIGHBOR_COUNT: return s->d_nbr_cnt.p;
        case ECMGPU_STATUS: return s->d_status.p;
        case ECMGPU_REPLAN_PENDING: return s->d_replan_pending.p;
        default: return nullptr;
    }
}

int check_range(ecmgpu_sim* s, int which, int first, int count) {
    if (!s) return ECMGPU_ERR_INVALID;
    if (elem_size(which) == 0) return fail(s, ECMGPU_ERR_INVALID, "unknown array selector");
    if (first < 0 || count < 0 || first + count > s->prm.max_agents) return fail(s, ECMGPU_ERR_INVALID, "slot range out of bounds");
    return ECMGPU_OK;
}

}  // namespace

// =================================================================================================
extern "C" {

const char* ecmgpu_last_error(const ecmgpu_sim* sim) { return sim ? sim->err.c_str() : g_create_error.c_str(); }

int ecmgpu_create(const ecmgpu_params* params, ecmgpu_sim** out) {
    if (!params || !out) return fail(nullptr, ECMGPU_ERR_INVALID, "ecmgpu_create: null argument");
    *out = nullptr;
    if (params->max_agents <= 0 || !(params->step > 0.0f)) return fail(nullptr, ECMGPU_ERR_INVALID, "ecmgpu_create: max_agents and step must be positive");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, ECMGPU_ERR_CUDA, std::string("no CUDA device (this library has no CPU fallback): ") + cudaGetErrorString(e));
    if (params->device < 0 || params->device >= ndev) return fail(nullptr, ECMGPU_ERR_INVALID, "ecmgpu_create: bad device ordinal");
    e = cudaSetDevice(params->device);
    if (e != cudaSuccess) return fail(nullptr, ECMGPU_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    ecmgpu_sim* s = new ecmgpu_sim();
    s->prm = *params;
    const size_t n = (size_t)params->max_agents;
    // payload capacity + room for aligning every path start to a block of 8 points
    const size_t pool = (params->path_pool_points > 0 ? (size_t)params->path_pool_points : 16 * n) + 8 * n + 8;
    auto bad = [&](cudaError_t ce, const char* what) {
        fail(nullptr, ECMGPU_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(ce));
        ecmgpu_destroy(s);
        return ECMGPU_ERR_CUDA;
    };
#define TRY_ALLOC(x) do { cudaError_t ce = (x); if (ce != cudaSuccess) return bad(ce, #x); } while (0)
    TRY_ALLOC(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    TRY_ALLOC(s->d_pos.alloc(n)); TRY_ALLOC(s->d_vel.alloc(n)); TRY_ALLOC(s->d_prefvel.alloc(n));
    TRY_ALLOC(s->d_attraction.alloc(n)); TRY_ALLOC(s->d_force.alloc(n));
    TRY_ALLOC(s->d_radius.alloc(n)); TRY_ALLOC(s->d_speed.alloc(n));
    TRY_ALLOC(s->d_active.alloc(n)); TRY_ALLOC(s->d_replan_pending.alloc(n));
    TRY_ALLOC(s->d_status.alloc(n)); TRY_ALLOC(s->d_cell.alloc(n));
    TRY_ALLOC(s->d_nbr.alloc(5 * n)); TRY_ALLOC(s->d_nbr_cnt.alloc(n));
    TRY_ALLOC(s->d_path_hdr.alloc(n)); TRY_ALLOC(s->d_path_pool.alloc(pool)); TRY_ALLOC(s->d_path_bbox.alloc(pool / 8 + 1));
    TRY_ALLOC(s->d_key.alloc(n)); TRY_ALLOC(s->d_rank.alloc(n));
    TRY_ALLOC(s->d_s_pos.alloc(n)); TRY_ALLOC(s->d_s_vel.alloc(n)); TRY_ALLOC(s->d_s_pref.alloc(n));
    TRY_ALLOC(s->d_s_rad.alloc(n)); TRY_ALLOC(s->d_s_spd.alloc(n)); TRY_ALLOC(s->d_s_slot.alloc(n));
    TRY_ALLOC(s->d_s_alive.alloc(n)); TRY_ALLOC(s->d_fb_list.alloc(n)); TRY_ALLOC(s->d_s_ghost.alloc(n));
    s->lp3d_cap = (int)n;
    TRY_ALLOC(s->d_lp3d_hdr.alloc(s->lp3d_cap)); TRY_ALLOC(s->d_lp3d_out.alloc(s->lp3d_cap));
    TRY_ALLOC(s->d_lp3d_cs.alloc((size_t)s->lp3d_cap * kMaxCons));
    TRY_ALLOC(s->d_ev_replan.alloc(n)); TRY_ALLOC(s->d_ev_destroyed.alloc(n));
    TRY_ALLOC(s->d_counters.alloc(C_COUNT));
    TRY_ALLOC(s->d_int_of.alloc(n)); TRY_ALLOC(s->d_ext_of.alloc(n));
    TRY_ALLOC(cudaMemsetAsync(s->d_pos.p, 0, 8 * n, s->stream)); TRY_ALLOC(cudaMemsetAsync(s->d_vel.p, 0, 8 * n, s->stream));
    TRY_ALLOC(cudaMemsetAsync(s->d_prefvel.p, 0, 8 * n, s->stream)); TRY_ALLOC(cudaMemsetAsync(s->d_attraction.p, 0, 8 * n, s->stream));
    TRY_ALLOC(cudaMemsetAsync(s->d_force.p, 0, 8 * n, s->stream));
    TRY_ALLOC(cudaMemsetAsync(s->d_radius.p, 0, 4 * n, s->stream)); TRY_ALLOC(cudaMemsetAsync(s->d_speed.p, 0, 4 * n, s->stream));
    TRY_ALLOC(cudaMemsetAsync(s->d_active.p, 0, n, s->stream)); TRY_ALLOC(cudaMemsetAsync(s->d_replan_pending.p, 0, n, s->stream));
    TRY_ALLOC(cudaMemsetAsync(s->d_status.p, 0, 4 * n, s->stream)); TRY_ALLOC(cudaMemsetAsync(s->d_cell.p, 0xff, 4 * n, s->stream));
    TRY_ALLOC(cudaMemsetAsync(s->d_nbr.p, 0xff, 20 * n, s->stream)); TRY_ALLOC(cudaMemsetAsync(s->d_nbr_cnt.p, 0, 4 * n, s->stream));
    TRY_ALLOC(cudaMemsetAsync(s->d_path_hdr.p, 0, sizeof(PathHdr) * n, s->stream));
    TRY_ALLOC(cudaMemsetAsync(s->d_counters.p, 0, sizeof(unsigned long long) * C_COUNT, s->stream));
    for (auto& ev : s->ev) TRY_ALLOC(cudaEventCreate(&ev));
    for (auto& ev : s->marks) TRY_ALLOC(cudaEventCreate(&ev));
    TRY_ALLOC(cudaStreamSynchronize(s->stream));
#undef TRY_ALLOC
    s->h_path_hdr.assign(n, PathHdr{0, 0, 0.0f, 0.0f});
    s->h_int_of.resize(n);
    s->h_ext_of.resize(n);
    for (size_t i = 0; i < n; i++) s->h_int_of[i] = s->h_ext_of[i] = (int)i;
    if (const char* e = getenv("ECMGPU_COHERENT")) s->coherent = atoi(e) != 0;
    if (const char* e = getenv("ECMGPU_GRAPH")) s->use_graph = atoi(e) != 0;
    if (const char* e = getenv("ECMGPU_PROFILE_GRAPH")) s->profile_in_graph = atoi(e) != 0;
    if (const char* e = getenv("ECMGPU_IO_DIRECT")) s->io.direct_ok = atoi(e) != 0;
    if (const char* e = getenv("ECMGPU_COMPACT")) s->compact = atoi(e) != 0;
    s->h_path_pool.reserve(std::min<size_t>(pool, 1 << 20));
    *out = s;
    return ECMGPU_OK;
}

void ecmgpu_destroy(ecmgpu_sim* s) {
    if (!s) return;
    if (s->stream) cudaStreamSynchronize(s->stream);
    for (auto& ev : s->ev) if (ev) cudaEventDestroy(ev);
    for (auto& ev : s->marks) if (ev) cudaEventDestroy(ev);
    for (auto& ev : s->pl_ev) if (ev) cudaEventDestroy(ev);
    s->d_vert_xy.free(); s->d_edge_cl.free(); s->d_obst_xy.free(); s->d_edge_v.free();
    s->d_obst_next.free(); s->d_obst_prev.free(); s->d_obst_convex.free(); s->d_obst_dir.free();
    s->d_bin_cell_start.free(); s->d_bin_cell_items.free(); s->d_bin_obst_start.free(); s->d_bin_obst_items.free();
    s->d_pos.free(); s->d_vel.free(); s->d_prefvel.free(); s->d_attraction.free(); s->d_force.free();
    s->d_radius.free(); s->d_speed.free(); s->d_active.free(); s->d_replan_pending.free(); s->d_status.free();
    s->d_cell.free(); s->d_nbr.free(); s->d_nbr_cnt.free(); s->d_path_hdr.free(); s->d_path_pool.free(); s->d_path_bbox.free();
    s->d_key.free(); s->d_rank.free(); s->d_cell_count.free(); s->d_scan_state.free(); s->d_scan_ctl.free(); s->d_s_slot.free();
    s->d_lp3d_hdr.free(); s->d_lp3d_out.free(); s->d_lp3d_cs.free();
    s->d_fb_list.free(); s->d_ev_replan.free(); s->d_ev_destroyed.free(); s->d_s_pos.free(); s->d_s_vel.free();
    s->d_s_pref.free(); s->d_s_rad.free(); s->d_s_spd.free(); s->d_s_alive.free(); s->d_counters.free();
    s->d_walk.free(); s->d_walk_n.free(); s->d_in_walk.free();
    s->d_int_of.free(); s->d_ext_of.free(); s->d_xfer.free();
    kd_free(s);
    plan_free(s);
    s->d_vert_clear.free(); s->d_vert_he.free(); s->d_he_next.free();
    comm_teardown(s);
    for (int g = 0; g < 2; g++) {
        if (s->graph_exec[g]) cudaGraphExecDestroy(s->graph_exec[g]);
        if (s->graph[g]) cudaGraphDestroy(s->graph[g]);
    }
    if (s->io.ready) {
        cudaStreamSynchronize(s->io.s_in);
        cudaStreamSynchronize(s->io.s_out);
        for (int b = 0; b < 2; b++) {
            cudaFree(s->io.in_buf[b]); cudaFree(s->io.out_buf[b]);
            cudaEventDestroy(s->io.in_done[b]); cudaEventDestroy(s->io.in_consumed[b]); cudaEventDestroy(s->io.tick_done[b]); cudaEventDestroy(s->io.out_done[b]);
        }
        for (int k = 0; k < s->io.kTickets; k++) {
            if (s->io.ticket_done[k]) cudaEventDestroy(s->io.ticket_done[k]);
        }
        cudaStreamDestroy(s->io.s_in);
        cudaStreamDestroy(s->io.s_out);
    }
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

int ecmgpu_set_ecm(ecmgpu_sim* s, const float bbox[4], int nV, const float* vert_xy, const float* vert_clear, int nE,
                   const int* edge_v, const float* edge_cl) {
    if (!s) return ECMGPU_ERR_INVALID;
    if (!bbox || nV <= 0 || nE <= 0 || !vert_xy || !edge_v || !edge_cl) return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_set_ecm: bad arguments");
    // vertex clearance is used by the planners only (the host's, and ecmgpu_plan_paths)
    for (int e = 0; e < 2 * nE; e++)
        if (edge_v[e] < 0 || edge_v[e] >= nV) return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_set_ecm: edge vertex index out of range");
    CUDA_TRY(s, cudaSetDevice(s->prm.device));
    memcpy(s->bbox, bbox, sizeof(float) * 4);
    s->n_vertices = nV;
    s->n_edges = nE;
    s->h_vert_xy.assign(vert_xy, vert_xy + 2 * nV);
    s->h_edge_v.assign(edge_v, edge_v + 2 * nE);
    s->h_edge_cl.assign(edge_cl, edge_cl + 8 * nE);
    CUDA_TRY(s, s->d_vert_xy.alloc(nV));
    CUDA_TRY(s, s->d_edge_v.alloc(nE));
    CUDA_TRY(s, s->d_edge_cl.alloc(4 * (size_t)nE));
    CUDA_TRY(s, cudaMemcpyAsync(s->d_vert_xy.p, vert_xy, sizeof(float) * 2 * nV, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(s, cudaMemcpyAsync(s->d_edge_v.p, edge_v, sizeof(int) * 2 * nE, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(s, cudaMemcpyAsync(s->d_edge_cl.p, edge_cl, sizeof(float) * 8 * nE, cudaMemcpyHostToDevice, s->stream));
    s->d_vert_clear.free();
    if (vert_clear) {
        CUDA_TRY(s, s->d_vert_clear.alloc(nV));
        CUDA_TRY(s, cudaMemcpyAsync(s->d_vert_clear.p, vert_clear, sizeof(float) * nV, cudaMemcpyHostToDevice, s->stream));
    }
    s->have_topology = false;  // belongs to the previous graph
    plan_free(s);
    CUDA_TRY(s, cudaStreamSynchronize(s->stream));
    s->have_ecm = true;
    s->bins_dirty = true;
    s->grid_dirty = true;
    return ECMGPU_OK;
}

int ecmgpu_set_obstacles(ecmgpu_sim* s, int n, const float* xy, const int* next, const int* prev, const uint8_t* convex) {
    if (!s) return ECMGPU_ERR_INVALID;
    if (n < 0 || (n > 0 && (!xy || !next || !prev || !convex))) return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_set_obstacles: bad arguments");
    for (int i = 0; i < n; i++)
        if (next[i] < 0 || next[i] >= n || prev[i] < 0 || prev[i] >= n) return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_set_obstacles: link out of range");
    CUDA_TRY(s, cudaSetDevice(s->prm.device));
    s->n_obst = n;
    s->h_obst_xy.assign(xy, xy + 2 * (size_t)n);
    s->h_obst_next.assign(next, next + n);
    CUDA_TRY(s, s->d_obst_xy.alloc(std::max(n, 1)));
    CUDA_TRY(s, s->d_obst_next.alloc(std::max(n, 1)));
    CUDA_TRY(s, s->d_obst_prev.alloc(std::max(n, 1)));
    CUDA_TRY(s, s->d_obst_convex.alloc(std::max(n, 1)));
    CUDA_TRY(s, s->d_obst_dir.alloc(std::max(n, 1)));
    std::vector<float2> dir(std::max(n, 1), make_float2(0.0f, 0.0f));
    for (int i = 0; i < n; i++) {  // Vec2::Normalize (ECMDataTypes.h:45-52) in IEEE float, no contraction
        volatile float dx = xy[2 * next[i]] - xy[2 * i], dy = xy[2 * next[i] + 1] - xy[2 * i + 1];
        volatile float xx = dx * dx, yy = dy * dy;
        volatile float l = sqrtf(xx + yy);
        dir[i] = l == 0.0f ? make_float2(dx, dy) : make_float2(dx / l, dy / l);
    }
    if (n > 0) {
        CUDA_TRY(s, cudaMemcpyAsync(s->d_obst_xy.p, xy, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(s, cudaMemcpyAsync(s->d_obst_next.p, next, sizeof(int) * n, cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(s, cudaMemcpyAsync(s->d_obst_prev.p, prev, sizeof(int) * n, cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(s, cudaMemcpyAsync(s->d_obst_convex.p, convex, n, cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(s, cudaMemcpyAsync(s->d_obst_dir.p, dir.data(), sizeof(float2) * n, cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(s, cudaStreamSynchronize(s->stream));
    }
    s->have_obst = true;
    s->bins_dirty = true;
    return ECMGPU_OK;
}

int ecmgpu_bulk_load(ecmgpu_sim* s, int n, const int* slots, const float* pos_xy, const float* radius, const float* speed,
                     const int* path_off, const float* path_xy) {
    if (!s) return ECMGPU_ERR_INVALID;
    if (n < 0 || !pos_xy || !radius || !speed || !path_off || !path_xy) return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_bulk_load: bad arguments");
    CUDA_TRY(s, cudaSetDevice(s->prm.device));
    int hi = s->n_slots;
    for (int i = 0; i < n; i++) {
        int slot = slots ? slots[i] : i;
        if (slot < 0 || slot >= s->prm.max_agents) return fail(s, ECMGPU_ERR_CAPACITY, "ecmgpu_bulk_load: slot out of range");
        if (path_off[i + 1] - path_off[i] < 1) return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_bulk_load: empty path");
        hi = std::max(hi, slot + 1);
    }
    if (n == 0) return ECMGPU_OK;
    s->n_slots = hi;
    s->io.owned_confirmed = -1;
    s->walk_dirty = true;
    // ---- spatial renumbering: the internal indices the loaded slots hold are handed out again in spatial order.  The
    // key uses nothing but the loaded positions, so every rank of a strips run (each loads the same crowd) derives the
    // same mapping: the halo / migrant messages can name agents by internal index.
    std::vector<int> order(n);  // order[r] = input position of the r-th agent in internal order
    for (int i = 0; i < n; i++) order[i] = i;
    std::vector<int> ids(n);    // ids[r] = internal index of that agent (ascending)
    for (int i = 0; i < n; i++) ids[i] = s->h_int_of[slots ? slots[i] : i];
    {
        std::vector<int> chk(ids);
        std::sort(chk.begin(), chk.end());
        if (std::adjacent_find(chk.begin(), chk.end()) != chk.end()) return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_bulk_load: a slot is listed twice");
    }
    const bool remap = s->coherent && n >= 64 && s->neighbor_mode == ECMGPU_NEIGHBORS_EXACT;
    if (remap) {
        double x0 = pos_xy[0], y0 = pos_xy[1], x1 = x0;
        for (int i = 1; i < n; i++) {
            x0 = std::min<double>(x0, pos_xy[2 * i]); x1 = std::max<double>(x1, pos_xy[2 * i]);
            y0 = std::min<double>(y0, pos_xy[2 * i + 1]);
        }
        const double cell = 4.0;
        const long long gw = (long long)std::floor((x1 - x0) / cell) + 1;
        std::vector<long long> key(n);
        for (int i = 0; i < n; i++) {
            const double fx = (pos_xy[2 * i] - x0) / cell, fy = (pos_xy[2 * i + 1] - y0) / cell;
            key[i] = (fx == fx && fy == fy) ? (long long)std::floor(fy) * gw + (long long)std::floor(fx) : -1;
        }
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key[a] < key[b]; });
        std::sort(ids.begin(), ids.end());
        for (int r = 0; r < n; r++) {
            const int slot = slots ? slots[order[r]] : order[r];
            s->h_int_of[slot] = ids[r];
            s->h_ext_of[ids[r]] = slot;
        }
        bool ident = true;
        for (int i = 0; i < s->n_slots && ident; i++) ident = s->h_int_of[i] == i;
        s->perm_identity = ident;
        CUDA_TRY(s, cudaMemcpyAsync(s->d_int_of.p, s->h_int_of.data(), sizeof(int) * (size_t)s->prm.max_agents, cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(s, cudaMemcpyAsync(s->d_ext_of.p, s->h_ext_of.data(), sizeof(int) * (size_t)s->prm.max_agents, cudaMemcpyHostToDevice, s->stream));
        s->config_epoch++;  // the captured tick holds the ext_of pointer (or its absence)
    }
    // ---- paths (host mirror indexed internally) and the obstacle range
    for (int r = 0; r < n; r++) {
        const int i = order[r];
        int rc = upload_path(s, ids[r], path_xy + 2 * (size_t)path_off[i], path_off[i + 1] - path_off[i]);
        if (rc) return rc;
        s->tracked_range = std::max(s->tracked_range, kLookAhead * speed[i] + radius[i]);
    }
    // ---- components: staged in internal order when the indices are one run, otherwise element by element
    bool contiguous = true;
    for (int r = 1; r < n && contiguous; r++) contiguous = ids[r] == ids[0] + r;
    const int first = ids[0];
    std::vector<unsigned char> ones(n, 1);
    if (contiguous) {
        std::vector<float2> hp(n);
        std::vector<float> hr(n), hs(n);
        for (int r = 0; r < n; r++) {
            const int i = order[r];
            hp[r] = make_float2(pos_xy[2 * i], pos_xy[2 * i + 1]);
            hr[r] = radius[i];
            hs[r] = speed[i];
        }
        CUDA_TRY(s, cudaMemcpyAsync(s->d_pos.p + first, hp.data(), sizeof(float2) * n, cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(s, cudaMemcpyAsync(s->d_radius.p + first, hr.data(), sizeof(float) * n, cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(s, cudaMemcpyAsync(s->d_speed.p + first, hs.data(), sizeof(float) * n, cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(s, cudaMemcpyAsync(s->d_active.p + first, ones.data(), n, cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(s, cudaMemsetAsync(s->d_vel.p + first, 0, sizeof(float2) * n, s->stream));
        CUDA_TRY(s, cudaMemsetAsync(s->d_prefvel.p + first, 0, sizeof(float2) * n, s->stream));
        CUDA_TRY(s, cudaMemsetAsync(s->d_force.p + first, 0, sizeof(float2) * n, s->stream));
        CUDA_TRY(s, cudaMemsetAsync(s->d_attraction.p + first, 0, sizeof(float2) * n, s->stream));
        CUDA_TRY(s, cudaMemsetAsync(s->d_replan_pending.p + first, 0, n, s->stream));
        { int rc = flush_pool(s); if (rc) return rc; }
        CUDA_TRY(s, cudaMemcpyAsync(s->d_path_hdr.p + first, s->h_path_hdr.data() + first, sizeof(PathHdr) * n, cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(s, cudaStreamSynchronize(s->stream));  // the staging vectors die here
    } else {
        for (int r = 0; r < n; r++) {
            const int i = order[r], a = ids[r];
            CUDA_TRY(s, cudaMemcpyAsync(s->d_pos.p + a, pos_xy + 2 * i, sizeof(float2), cudaMemcpyHostToDevice, s->stream));
            CUDA_TRY(s, cudaMemcpyAsync(s->d_radius.p + a, radius + i, sizeof(float), cudaMemcpyHostToDevice, s->stream));
            CUDA_TRY(s, cudaMemcpyAsync(s->d_speed.p + a, speed + i, sizeof(float), cudaMemcpyHostToDevice, s->stream));
            CUDA_TRY(s, cudaMemcpyAsync(s->d_active.p + a, ones.data(), 1, cudaMemcpyHostToDevice, s->stream));
            CUDA_TRY(s, cudaMemsetAsync(s->d_vel.p + a, 0, sizeof(float2), s->stream));
            CUDA_TRY(s, cudaMemsetAsync(s->d_prefvel.p + a, 0, sizeof(float2), s->stream));
            CUDA_TRY(s, cudaMemsetAsync(s->d_force.p + a, 0, sizeof(float2), s->stream));
            CUDA_TRY(s, cudaMemsetAsync(s->d_attraction.p + a, 0, sizeof(float2), s->stream));
            CUDA_TRY(s, cudaMemsetAsync(s->d_replan_pending.p + a, 0, 1, s->stream));
        }
        { int rc = flush_pool(s); if (rc) return rc; }
        for (int r = 0; r < n; r++)
            CUDA_TRY(s, cudaMemcpyAsync(s->d_path_hdr.p + ids[r], &s->h_path_hdr[ids[r]], sizeof(PathHdr), cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(s, cudaStreamSynchronize(s->stream));
    }
    return ECMGPU_OK;
}

int ecmgpu_spawn(ecmgpu_sim* s, int slot, float x, float y, float radius, float speed, const float* path_xy, int n_points) {
    const int off[2] = {0, n_points};
    const float pos[2] = {x, y};
    return ecmgpu_bulk_load(s, 1, &slot, pos, &radius, &speed, off, path_xy);
}

int ecmgpu_set_path(ecmgpu_sim* s, int slot, const float* path_xy, int n_points) {
    if (!s) return ECMGPU_ERR_INVALID;
    if (slot < 0 || slot >= s->n_slots || !path_xy) return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_set_path: bad arguments");
    CUDA_TRY(s, cudaSetDevice(s->prm.device));
    const int a = s->h_int_of[slot];
    int rc = upload_path(s, a, path_xy, n_points);
    if (rc) return rc;
    const unsigned char zero = 0;
    { int rc2 = flush_pool(s); if (rc2) return rc2; }
    CUDA_TRY(s, cudaMemcpyAsync(s->d_path_hdr.p + a, &s->h_path_hdr[a], sizeof(PathHdr), cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(s, cudaMemcpyAsync(s->d_replan_pending.p + a, &zero, 1, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(s, cudaStreamSynchronize(s->stream));
    return ECMGPU_OK;
}

int ecmgpu_destroy_agent(ecmgpu_sim* s, int slot) {
    if (!s) return ECMGPU_ERR_INVALID;
    if (slot < 0 || slot >= s->prm.max_agents) return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_destroy_agent: bad slot");
    CUDA_TRY(s, cudaSetDevice(s->prm.device));
    CUDA_TRY(s, cudaMemsetAsync(s->d_active.p + s->h_int_of[slot], 0, 1, s->stream));
    return ECMGPU_OK;
}

// Phase boundary for ecmgpu_last_tick_ms.  Inside a stream capture a plain cudaEventRecord only orders captured work;
// the external flag makes it an event-record node that stamps the event at every replay of the graph.
cudaError_t record_phase_event(ecmgpu_sim* s, int i) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    cudaError_t e = cudaStreamIsCapturing(s->stream, &st);
    if (e != cudaSuccess) return e;
    if (st == cudaStreamCaptureStatusActive) return cudaEventRecordWithFlags(s->ev[i], s->stream, cudaEventRecordExternal);
    return cudaEventRecord(s->ev[i], s->stream);
}

int ecmgpu_update_phase(ecmgpu_sim* s, int phase) {
    if (!s) return ECMGPU_ERR_INVALID;
    if (phase < 0 || phase > 2) return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_update_phase: phase must be 0, 1 or 2");
    CUDA_TRY(s, cudaSetDevice(s->prm.device));
    if (s->n_slots == 0) { if (phase == 2) s->ticks++; return ECMGPU_OK; }
    int rc = ensure_ready(s);
    if (rc) return rc;
    TickView t = make_view(s);
    if (phase == 0) {
        if (s->profiling) CUDA_TRY(s, record_phase_event(s, 0));
        return s->strips_on ? enqueue_pack(s, t) : ECMGPU_OK;
    }
    if (phase == 1) return s->strips_on ? enqueue_exchange(s, t) : ECMGPU_OK;
    rc = enqueue_grid_build(s, t);
    if (rc) return rc;
    if (s->profiling) CUDA_TRY(s, record_phase_event(s, 1));
    const int nb = div_up(launch_slots(s) + (s->strips_on ? 2 * s->cap_halo + s->cap_self : 0), 128);
    if (s->neighbor_mode == ECMGPU_NEIGHBORS_KDTREE) {
        if (s->strips_on) return fail(s, ECMGPU_ERR_INVALID, "the KD-tree neighbour mode runs on a single handle (no strips)");
        k_attract<<<nb, 128, 0, s->stream>>>(t);
        s->launches++;
        if (s->profiling) CUDA_TRY(s, record_phase_event(s, 2));
        rc = enqueue_kd_orca(s, t, nb * 128);
        if (rc) return rc;
        s->launches -= 2;  // k_attract and k_orca_kd are counted above (3 are added below)
    } else if (s->compact && s->strips_on) {  // one resident wave over the row tiles that exist (tick.cuh)
        k_attract_tiles<<<kSMs * ECM_ATTRACT_MINBLOCKS, 128, 0, s->stream>>>(t);
        if (s->profiling) CUDA_TRY(s, record_phase_event(s, 2));
        k_orca_tiles<<<kSMs * ECM_ORCA_MINBLOCKS, ECM_ORCA_BLOCK, 0, s->stream>>>(t);
    } else {
        k_attract<<<nb, 128, 0, s->stream>>>(t);
        if (s->profiling) CUDA_TRY(s, record_phase_event(s, 2));
        k_orca<<<div_up(nb * 128, ECM_ORCA_BLOCK), ECM_ORCA_BLOCK, 0, s->stream>>>(t);
    }
    if (s->profiling) CUDA_TRY(s, record_phase_event(s, 4));  // k_orca | k_fallback (ecmgpu_last_tick_phases)
    k_fallback<<<kSMs * ECM_FALLBACK_CTAS, 128, 0, s->stream>>>(t, 0);
    if (s->profiling) { CUDA_TRY(s, record_phase_event(s, 3)); s->ev_valid = true; }
    s->launches += 3;
    s->ticks++;
    CUDA_TRY(s, cudaGetLastError());
    return ECMGPU_OK;
}

int ecmgpu_update(ecmgpu_sim* s) {
    if (!s) return ECMGPU_ERR_INVALID;
    if (s->local_transport && s->n_ranks > 1)
        return fail(s, ECMGPU_ERR_INVALID, "in-process strips: drive all handles with ecmgpu_update_phase(0), (1), (2)");
    // NCCL send/recv are kept out of graph capture (capturing them hung on 4 x B200 with NCCL 2.28): with the
    // NCCL transport the tick is submitted launch by launch.  The peer transport is plain kernels and memsets.
    const bool nccl_tick = s->strips_on && !s->local_transport && s->n_ranks > 1 && !s->p2p;
    // the KD-tree mode submits a library sort per tree level: launch by launch as well
    // (profiling: the phase events are captured as event-record nodes, so the phases are timed inside the very tick the
    // unprofiled run replays, not in a launch-by-launch one with its host gaps; ECMGPU_PROFILE_GRAPH=0: launch by launch)
    if (s->use_graph && !nccl_tick && (!s->profiling || s->profile_in_graph) && s->n_slots > 0 && s->neighbor_mode == ECMGPU_NEIGHBORS_EXACT) {
        CUDA_TRY(s, cudaSetDevice(s->prm.device));
        int rc = ensure_ready(s);  // host-side (re)builds happen outside the capture
        if (rc) return rc;
        rc = ensure_walk(s);
        if (rc) return rc;
        const int g = (int)(s->comm_seq & 1u);
        if (!s->graph_exec[g] || s->graph_epoch[g] != s->config_epoch || s->graph_n_slots[g] != launch_slots(s)) {
            if (s->graph_exec[g]) { cudaGraphExecDestroy(s->graph_exec[g]); s->graph_exec[g] = nullptr; }
            if (s->graph[g]) { cudaGraphDestroy(s->graph[g]); s->graph[g] = nullptr; }
            const uint64_t l0 = s->launches, t0 = s->ticks;
            const unsigned seq0 = s->comm_seq;
            CUDA_TRY(s, cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeRelaxed));
            for (int phase = 0; phase < 3 && rc == ECMGPU_OK; phase++) rc = ecmgpu_update_phase(s, phase);
            cudaError_t ce = cudaStreamEndCapture(s->stream, &s->graph[g]);
            s->graph_launches[g] = s->launches - l0;
            s->launches = l0;
            s->ticks = t0;
            s->comm_seq = seq0;
            if (rc) return rc;
            if (ce != cudaSuccess) return fail(s, ECMGPU_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
            CUDA_TRY(s, cudaGraphInstantiate(&s->graph_exec[g], s->graph[g], 0));
            s->graph_epoch[g] = s->config_epoch;
            s->graph_n_slots[g] = launch_slots(s);
        }
        CUDA_TRY(s, cudaGraphLaunch(s->graph_exec[g], s->stream));
        s->launches += s->graph_launches[g];
        s->ticks++;
        if (s->strips_on && s->n_ranks > 1) {  // what enqueue_pack / enqueue_exchange do to the host-side state
            s->cur_gen = s->comm_seq & 1u;
            s->comm_seq++;
        }
        return ECMGPU_OK;
    }
    for (int phase = 0; phase < 3; phase++) {
        int rc = ecmgpu_update_phase(s, phase);
        if (rc) return rc;
    }
    return ECMGPU_OK;
}

int ecmgpu_sync(ecmgpu_sim* s) {
    if (!s) return ECMGPU_ERR_INVALID;
    CUDA_TRY(s, cudaSetDevice(s->prm.device));
    CUDA_TRY(s, cudaStreamSynchronize(s->stream));
    return ECMGPU_OK;
}

int ecmgpu_poll_events(ecmgpu_sim* s, int* replan_slots, int replan_cap, int* n_replans, int* destroyed_slots, int destroyed_cap,
                       int* n_destroyed) {
    if (!s) return ECMGPU_ERR_INVALID;
    CUDA_TRY(s, cudaSetDevice(s->prm.device));
    unsigned long long c[2] = {0, 0};
    CUDA_TRY(s, cudaMemcpyAsync(c, s->d_counters.p, sizeof(c), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(s, cudaStreamSynchronize(s->stream));
    // the lists hold max_agents entries each; what a kernel could not store it dropped and counted
    const unsigned long long cap = (unsigned long long)s->d_ev_destroyed.n;
    const bool overflow = c[C_REPLAN_N] > cap || c[C_DESTROYED_N] > cap;
    const int nr = (int)std::min(c[C_REPLAN_N], cap), nd = (int)std::min(c[C_DESTROYED_N], cap);
    if (n_replans) *n_replans = nr;
    if (n_destroyed) *n_destroyed = nd;
    if (replan_slots && nr > 0) {
        if (replan_cap < nr) return fail(s, ECMGPU_ERR_CAPACITY, "ecmgpu_poll_events: replan buffer too small");
        CUDA_TRY(s, cudaMemcpyAsync(replan_slots, s->d_ev_replan.p, sizeof(int) * nr, cudaMemcpyDeviceToHost, s->stream));
    }
    if (destroyed_slots && nd > 0) {
        if (destroyed_cap < nd) return fail(s, ECMGPU_ERR_CAPACITY, "ecmgpu_poll_events: destroyed buffer too small");
        CUDA_TRY(s, cudaMemcpyAsync(destroyed_slots, s->d_ev_destroyed.p, sizeof(int) * nd, cudaMemcpyDeviceToHost, s->stream));
    }
    // drain only what was handed out
    if ((replan_slots || replan_cap == 0) && (destroyed_slots || destroyed_cap == 0) && (replan_slots || destroyed_slots)) {
        if (replan_slots) CUDA_TRY(s, cudaMemsetAsync(s->d_counters.p + C_REPLAN_N, 0, sizeof(unsigned long long), s->stream));
        if (destroyed_slots) CUDA_TRY(s, cudaMemsetAsync(s->d_counters.p + C_DESTROYED_N, 0, sizeof(unsigned long long), s->stream));
    }
    CUDA_TRY(s, cudaStreamSynchronize(s->stream));
    // the kernels name agents by internal index
    if (!s->perm_identity) {
        if (replan_slots) for (int k = 0; k < nr; k++) replan_slots[k] = s->h_ext_of[replan_slots[k]];
        if (destroyed_slots) for (int k = 0; k < nd; k++) destroyed_slots[k] = s->h_ext_of[destroyed_slots[k]];
    }
    // event order is arrival order of the atomics; report ascending slots like the reference's loops
    if (replan_slots && nr > 1) std::sort(replan_slots, replan_slots + nr);
    if (destroyed_slots && nd > 1) std::sort(destroyed_slots, destroyed_slots + nd);
    if (overflow)
        return fail(s, ECMGPU_ERR_CAPACITY, "ecmgpu_poll_events: an event list overflowed (more than max_agents events between two polls); "
                    "the excess was dropped - read ECMGPU_ACTIVE / ECMGPU_REPLAN_PENDING to resynchronise");
    return ECMGPU_OK;
}

// dst[i] = src[int_of[first + i]] (gather) or the reverse (scatter) for one of the per-agent arrays; `es` = element size
static void launch_remap(ecmgpu_sim* s, bool gather, size_t es, int count, int first, const void* src, void* dst) {
    const int nb = div_up(count, 256);
    const int* m = s->d_int_of.p;
#define ECM_REMAP(T)                                                                                              \
    do {                                                                                                          \
        if (gather) k_gather_slots<T><<<nb, 256, 0, s->stream>>>(count, first, m, (const T*)src, (T*)dst);         \
        else k_scatter_slots<T><<<nb, 256, 0, s->stream>>>(count, first, m, (const T*)src, (T*)dst);               \
    } while (0)
    if (es == 1) ECM_REMAP(unsigned char);
    else if (es == 4) ECM_REMAP(unsigned);
    else if (es == 8) ECM_REMAP(float2);
    else ECM_REMAP(Int5);
#undef ECM_REMAP
    s->launches++;
}

static int xfer(ecmgpu_sim* s, int which, void* host, int first, int count, bool to_host, bool wait) {
    int rc = check_range(s, which, first, count);
    if (rc) return rc;
    if (!host) return fail(s, ECMGPU_ERR_INVALID, "null host buffer");
    CUDA_TRY(s, cudaSetDevice(s->prm.device));
    const size_t es = elem_size(which);
    if (s->perm_identity || count == 0) {
        char* dev = (char*)dev_array(s, which) + es * (size_t)first;
        if (to_host) CUDA_TRY(s, cudaMemcpyAsync(host, dev, es * (size_t)count, cudaMemcpyDeviceToHost, s->stream));
        else CUDA_TRY(s, cudaMemcpyAsync(dev, host, es * (size_t)count, cudaMemcpyHostToDevice, s->stream));
    } else {  // the arrays are indexed internally: stage the slot range through a gather / scatter kernel
        const size_t bytes = es * (size_t)count;
        if (s->d_xfer.n < bytes) {
            CUDA_TRY(s, cudaStreamSynchronize(s->stream));  // an earlier asynchronous transfer may still use the old staging
            CUDA_TRY(s, s->d_xfer.alloc(bytes + bytes / 2 + 1024));
        }
        if (to_host) {
            launch_remap(s, true, es, count, first, dev_array(s, which), s->d_xfer.p);
            CUDA_TRY(s, cudaMemcpyAsync(host, s->d_xfer.p, bytes, cudaMemcpyDeviceToHost, s->stream));
        } else {
            CUDA_TRY(s, cudaMemcpyAsync(s->d_xfer.p, host, bytes, cudaMemcpyHostToDevice, s->stream));
            launch_remap(s, false, es, count, first, s->d_xfer.p, dev_array(s, which));
        }
        CUDA_TRY(s, cudaGetLastError());
    }
    if (wait) CUDA_TRY(s, cudaStreamSynchronize(s->stream));
    if (!to_host && (which == ECMGPU_RADIUS || which == ECMGPU_SPEED) && count > 0) {
        // The obstacle lists of the static bins reach max(10 * speed + radius) (ORCA.cpp:27): a larger speed or radius
        // must widen them, or find_obstacles would silently miss segments.  Rare call: read everything back and re-track.
        const int m = std::max(s->n_slots, first + count);
        std::vector<float> spd(m), rad(m);
        CUDA_TRY(s, cudaMemcpyAsync(spd.data(), s->d_speed.p, sizeof(float) * m, cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(s, cudaMemcpyAsync(rad.data(), s->d_radius.p, sizeof(float) * m, cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(s, cudaStreamSynchronize(s->stream));
        for (int i = 0; i < m; i++) {
            const float r = kLookAhead * spd[i] + rad[i];
            if (r == r && r < std::numeric_limits<float>::infinity()) s->tracked_range = std::max(s->tracked_range, r);
        }
    }
    if (!to_host) s->n_slots = std::max(s->n_slots, which == ECMGPU_ACTIVE ? first + count : s->n_slots);
    if (!to_host && which == ECMGPU_ACTIVE) { s->io.owned_confirmed = -1; s->walk_dirty = true; }
    return ECMGPU_OK;
}
int ecmgpu_read(ecmgpu_sim* s, int which, void* dst, int first, int count) { return xfer(s, which, dst, first, count, true, true); }
int ecmgpu_write(ecmgpu_sim* s, int which, const void* src, int first, int count) { return xfer(s, which, (void*)src, first, count, false, true); }
int ecmgpu_read_async(ecmgpu_sim* s, int which, void* dst, int first, int count) { return xfer(s, which, dst, first, count, true, false); }
int ecmgpu_write_async(ecmgpu_sim* s, int which, const void* src, int first, int count) { return xfer(s, which, (void*)src, first, count, false, false); }

// One tick with host I/O, pipelined: upload (copy stream A) -> tick (main stream) -> download (copy
// stream B).  Consecutive calls overlap: while tick k computes, the inputs of k+1 are already on
// their way and the results of k-1 are still draining.  Device staging is double-buffered.
static int io_prepare(ecmgpu_sim* s) {
    auto& io = s->io;
    if (io.ready) return ECMGPU_OK;
    const size_t n = (size_t)s->prm.max_agents;
    CUDA_TRY(s, cudaStreamCreateWithFlags(&io.s_in, cudaStreamNonBlocking));
    CUDA_TRY(s, cudaStreamCreateWithFlags(&io.s_out, cudaStreamNonBlocking));
    for (int b = 0; b < 2; b++) {
        CUDA_TRY(s, cudaMalloc((void**)&io.in_buf[b], sizeof(ecmgpu_agent_rec) * n + 32));
        CUDA_TRY(s, cudaMalloc((void**)&io.out_buf[b], sizeof(ecmgpu_agent_rec) * n + 32));
        CUDA_TRY(s, cudaEventCreateWithFlags(&io.in_done[b], cudaEventDisableTiming));
        CUDA_TRY(s, cudaEventCreateWithFlags(&io.in_consumed[b], cudaEventDisableTiming));
        CUDA_TRY(s, cudaEventCreateWithFlags(&io.tick_done[b], cudaEventDisableTiming));
        CUDA_TRY(s, cudaEventCreateWithFlags(&io.out_done[b], cudaEventDisableTiming));
        CUDA_TRY(s, cudaEventRecord(io.in_consumed[b], s->stream));
        CUDA_TRY(s, cudaEventRecord(io.out_done[b], io.s_out));
    }
    for (int k = 0; k < io.kTickets; k++) CUDA_TRY(s, cudaEventCreateWithFlags(&io.ticket_done[k], cudaEventDisableTiming));
    io.ready = true;
    return ECMGPU_OK;
}

int ecmgpu_update_io(ecmgpu_sim* s, int count, const float* in_pos, const float* in_vel, float* out_pos, float* out_vel,
                     uint8_t* out_active, uint64_t* ticket) {
    if (!s) return ECMGPU_ERR_INVALID;
    if (count < 0 || count > s->prm.max_agents) return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_update_io: bad count");
    CUDA_TRY(s, cudaSetDevice(s->prm.device));
    int rc = io_prepare(s);
    if (rc) return rc;
    auto& io = s->io;
    const int b = (int)(io.calls & 1);
    const size_t c = (size_t)count, n8 = sizeof(float2) * (size_t)s->prm.max_agents;
    unsigned char *si = io.in_buf[b], *so = io.out_buf[b];
    // upload into staging b once the tick that last used it has consumed it
    CUDA_TRY(s, cudaStreamWaitEvent(io.s_in, io.in_consumed[b], 0));
    if (in_pos) CUDA_TRY(s, cudaMemcpyAsync(si, in_pos, sizeof(float2) * c, cudaMemcpyHostToDevice, io.s_in));
    if (in_vel) CUDA_TRY(s, cudaMemcpyAsync(si + n8, in_vel, sizeof(float2) * c, cudaMemcpyHostToDevice, io.s_in));
    CUDA_TRY(s, cudaEventRecord(io.in_done[b], io.s_in));
    // main stream: adopt the inputs, tick, publish the results
    CUDA_TRY(s, cudaStreamWaitEvent(s->stream, io.in_done[b], 0));
    // host arrays are indexed by slot, the device arrays internally (spatial renumbering): copy, or scatter through int_of
    if (s->perm_identity) {
        if (in_pos) CUDA_TRY(s, cudaMemcpyAsync(s->d_pos.p, si, sizeof(float2) * c, cudaMemcpyDeviceToDevice, s->stream));
        if (in_vel) CUDA_TRY(s, cudaMemcpyAsync(s->d_vel.p, si + n8, sizeof(float2) * c, cudaMemcpyDeviceToDevice, s->stream));
    } else if (count > 0 && (in_pos || in_vel)) {
        k_io_adopt<<<div_up(count, 256), 256, 0, s->stream>>>(count, s->d_int_of.p, in_pos ? (const float2*)si : nullptr,
                                                              in_vel ? (const float2*)(si + n8) : nullptr, s->d_pos.p, s->d_vel.p);
        s->launches++;
    }
    CUDA_TRY(s, cudaEventRecord(io.in_consumed[b], s->stream));
    rc = ecmgpu_update(s);
    if (rc) return rc;
    CUDA_TRY(s, cudaStreamWaitEvent(s->stream, io.out_done[b], 0));  // the download that last read staging b is over
    if (s->perm_identity) {
        if (out_pos) CUDA_TRY(s, cudaMemcpyAsync(so, s->d_pos.p, sizeof(float2) * c, cudaMemcpyDeviceToDevice, s->stream));
        if (out_vel) CUDA_TRY(s, cudaMemcpyAsync(so + n8, s->d_vel.p, sizeof(float2) * c, cudaMemcpyDeviceToDevice, s->stream));
        if (out_active) CUDA_TRY(s, cudaMemcpyAsync(so + 2 * n8, s->d_active.p, c, cudaMemcpyDeviceToDevice, s->stream));
    } else if (count > 0 && (out_pos || out_vel || out_active)) {
        k_io_publish<<<div_up(count, 256), 256, 0, s->stream>>>(count, s->d_int_of.p, s->d_pos.p, s->d_vel.p, s->d_active.p,
                                                                out_pos ? (float2*)so : nullptr, out_vel ? (float2*)(so + n8) : nullptr,
                                                                out_active ? so + 2 * n8 : nullptr);
        s->launches++;
    }
    CUDA_TRY(s, cudaEventRecord(io.tick_done[b], s->stream));
    // download
    CUDA_TRY(s, cudaStreamWaitEvent(io.s_out, io.tick_done[b], 0));
    if (out_pos) CUDA_TRY(s, cudaMemcpyAsync(out_pos, so, sizeof(float2) * c, cudaMemcpyDeviceToHost, io.s_out));
    if (out_vel) CUDA_TRY(s, cudaMemcpyAsync(out_vel, so + n8, sizeof(float2) * c, cudaMemcpyDeviceToHost, io.s_out));
    if (out_active) CUDA_TRY(s, cudaMemcpyAsync(out_active, so + 2 * n8, c, cudaMemcpyDeviceToHost, io.s_out));
    CUDA_TRY(s, cudaEventRecord(io.out_done[b], io.s_out));
    CUDA_TRY(s, cudaEventRecord(io.ticket_done[io.calls % io.kTickets], io.s_out));
    io.owned_count[io.calls % io.kTickets] = nullptr;
    if (ticket) *ticket = io.calls;
    io.calls++;
    return ECMGPU_OK;
}

// The same pipeline moving only the agents this handle owns, as (slot, position, velocity) records:
// with strips every rank transfers its share of the crowd instead of all slots.
int ecmgpu_update_io_owned(ecmgpu_sim* s, int n_in, const ecmgpu_agent_rec* in, ecmgpu_agent_rec* out, int out_cap, int32_t* out_count,
                           uint64_t* ticket) {
    if (!s) return ECMGPU_ERR_INVALID;
    if (n_in < 0 || n_in > s->prm.max_agents || (n_in > 0 && !in)) return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_update_io_owned: bad input records");
    if (!out || !out_count || out_cap < 0) return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_update_io_owned: bad output buffer");
    CUDA_TRY(s, cudaSetDevice(s->prm.device));
    int rc = io_prepare(s);
    if (rc) return rc;
    auto& io = s->io;
    const int b = (int)(io.calls & 1);
    static_assert(sizeof(AgentRec) == sizeof(ecmgpu_agent_rec), "record layout");
    AgentRec* si = (AgentRec*)io.in_buf[b];
    int* so_count = (int*)io.out_buf[b];
    AgentRec* so = (AgentRec*)(io.out_buf[b] + 32);
    CUDA_TRY(s, cudaStreamWaitEvent(io.s_in, io.in_consumed[b], 0));
    if (n_in) CUDA_TRY(s, cudaMemcpyAsync(si, in, sizeof(ecmgpu_agent_rec) * (size_t)n_in, cudaMemcpyHostToDevice, io.s_in));
    CUDA_TRY(s, cudaEventRecord(io.in_done[b], io.s_in));
    CUDA_TRY(s, cudaStreamWaitEvent(s->stream, io.in_done[b], 0));
    if (n_in) {
        k_apply_records<<<div_up(n_in, 256), 256, 0, s->stream>>>(n_in, si, s->prm.max_agents, s->d_active.p, s->d_pos.p, s->d_vel.p,
                                                                    s->perm_identity ? nullptr : s->d_int_of.p);
        s->launches++;
    }
    CUDA_TRY(s, cudaEventRecord(io.in_consumed[b], s->stream));
    rc = ecmgpu_update(s);
    if (rc) return rc;
    CUDA_TRY(s, cudaStreamWaitEvent(s->stream, io.out_done[b], 0));
    CUDA_TRY(s, cudaMemsetAsync(so_count, 0, sizeof(int), s->stream));
    // Where the records go.  The host cannot know this tick's count when the download is issued, so the staged copy is
    // sized from a bound (the last confirmed count plus room for migrants).  ECMGPU_IO_DIRECT=1 (opt-in) lets the collect
    // kernel store the records straight into `out` when that is pinned memory the device can address: exactly the owned
    // ones, no staging, no bound - but 20-byte records stored by SM threads cross PCIe as small writes and the kernel sits
    // on the tick's stream: measured 4.1 ms instead of 1.4 ms per tick end to end on the 4 M map on 8 GPUs
    // (profiles/r03p_bench_c4_4m_n8_direct_io.json).  Kept for hosts behind a coherent link.
    AgentRec* direct = nullptr;
    if (io.direct_ok && out_cap > 0) {
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, out) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer) direct = (AgentRec*)pa.devicePointer;
        else cudaGetLastError();  // an unregistered pointer is not an error of this call
    }
    if (s->n_slots > 0) {
        const StripView sv = make_strip_view(s);
        const int* ext = s->perm_identity ? nullptr : s->d_ext_of.p;
        AgentRec* dst = direct ? direct : so;
        const int dst_cap = direct ? out_cap : s->n_slots;
        if (sv.walk.list && !s->walk_dirty) k_collect_owned_walk<<<kSMs * 2, kCollectBlock, 0, s->stream>>>(sv.walk, s->d_active.p, s->d_pos.p, s->d_vel.p, dst, dst_cap, so_count, ext);
        else k_collect_owned<<<div_up(s->n_slots, kCollectBlock), kCollectBlock, 0, s->stream>>>(s->n_slots, s->d_active.p, s->d_pos.p, s->d_vel.p, dst, dst_cap, so_count, ext);
        s->launches++;
    }
    CUDA_TRY(s, cudaEventRecord(io.tick_done[b], s->stream));
    int copied = out_cap;
    CUDA_TRY(s, cudaStreamWaitEvent(io.s_out, io.tick_done[b], 0));
    CUDA_TRY(s, cudaMemcpyAsync(out_count, so_count, sizeof(int), cudaMemcpyDeviceToHost, io.s_out));
    if (!direct) {
        long long bound = s->n_slots;
        if (io.owned_confirmed >= 0) {
            const long long per_tick = s->strips_on ? 2ll * s->cap_migr : 0ll;
            // every tick since the confirmed one - through this call or plain ecmgpu_update - may have brought migrants in
            bound = std::min(bound, io.owned_confirmed + per_tick * (long long)(s->ticks - io.owned_confirmed_tick));
        }
        copied = (int)std::min<long long>(bound, out_cap);
        if (copied) CUDA_TRY(s, cudaMemcpyAsync(out, so, sizeof(ecmgpu_agent_rec) * (size_t)copied, cudaMemcpyDeviceToHost, io.s_out));
    }
    CUDA_TRY(s, cudaEventRecord(io.out_done[b], io.s_out));
    CUDA_TRY(s, cudaEventRecord(io.ticket_done[io.calls % io.kTickets], io.s_out));
    io.owned_count[io.calls % io.kTickets] = out_count;
    io.owned_copied[io.calls % io.kTickets] = copied;
    io.owned_tick[io.calls % io.kTickets] = s->ticks;
    if (ticket) *ticket = io.calls;
    io.calls++;
    return ECMGPU_OK;
}

int ecmgpu_io_wait(ecmgpu_sim* s, uint64_t ticket) {
    if (!s) return ECMGPU_ERR_INVALID;
    auto& io = s->io;
    if (!io.ready || ticket >= io.calls) return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_io_wait: unknown ticket");
    // a ticket slot that has been reused since: waiting for its latest user also covers the older call
    while (ticket + io.kTickets < io.calls) ticket += io.kTickets;
    CUDA_TRY(s, cudaSetDevice(s->prm.device));
    const int k = (int)(ticket % io.kTickets);
    CUDA_TRY(s, cudaEventSynchronize(io.ticket_done[k]));
    if (io.owned_count[k]) {
        const int have = *io.owned_count[k];
        if (io.owned_confirmed < 0 || ticket >= io.owned_confirmed_ticket) {
            io.owned_confirmed = have;
            io.owned_confirmed_ticket = ticket;
            io.owned_confirmed_tick = io.owned_tick[k];
        }
        if (have > io.owned_copied[k])
            return fail(s, ECMGPU_ERR_CAPACITY, "ecmgpu_update_io_owned: " + std::to_string(have) + " owned agents, room for " +
                        std::to_string(io.owned_copied[k]) + " records");
    }
    return ECMGPU_OK;
}

void* ecmgpu_alloc_pinned(uint64_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void ecmgpu_free_pinned(void* p) { if (p) cudaFreeHost(p); }

int ecmgpu_locate(ecmgpu_sim* s, int n, const float* xy, int* out_cell) {
    if (!s) return ECMGPU_ERR_INVALID;
    if (n < 0 || !xy || !out_cell) return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_locate: bad arguments");
    CUDA_TRY(s, cudaSetDevice(s->prm.device));
    if (!s->have_ecm) return fail(s, ECMGPU_ERR_INVALID, "no ECM: call ecmgpu_set_ecm first");
    if (s->bins_dirty) { int rc = build_bins(s); if (rc) return rc; }
    if (n == 0) return ECMGPU_OK;
    DevBuf<float2> dxy; DevBuf<int> dout;
    CUDA_TRY(s, dxy.alloc(n)); CUDA_TRY(s, dout.alloc(n));
    TickView t = make_view(s);
    CUDA_TRY(s, cudaMemcpyAsync(dxy.p, xy, sizeof(float2) * n, cudaMemcpyHostToDevice, s->stream));
    k_locate<<<div_up(n, 128), 128, 0, s->stream>>>(t.ecm, t.bins, n, dxy.p, dout.p);
    s->launches++;
    CUDA_TRY(s, cudaMemcpyAsync(out_cell, dout.p, sizeof(int) * n, cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(s, cudaStreamSynchronize(s->stream));
    dxy.free(); dout.free();
    return ECMGPU_OK;
}

int ecmgpu_retract(ecmgpu_sim* s, int n, const float* xy, uint8_t* out_ok, float* out_xy, int* out_edge) {
    if (!s) return ECMGPU_ERR_INVALID;
    if (n < 0 || !xy || !out_ok || !out_xy || !out_edge) return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_retract: bad arguments");
    CUDA_TRY(s, cudaSetDevice(s->prm.device));
    if (!s->have_ecm) return fail(s, ECMGPU_ERR_INVALID, "no ECM: call ecmgpu_set_ecm first");
    if (s->bins_dirty) { int rc = build_bins(s); if (rc) return rc; }
    if (n == 0) return ECMGPU_OK;
    DevBuf<float2> dxy, dout; DevBuf<int> dedge; DevBuf<unsigned char> dok;
    CUDA_TRY(s, dxy.alloc(n)); CUDA_TRY(s, dout.alloc(n)); CUDA_TRY(s, dedge.alloc(n)); CUDA_TRY(s, dok.alloc(n));
    TickView t = make_view(s);
    CUDA_TRY(s, cudaMemcpyAsync(dxy.p, xy, sizeof(float2) * n, cudaMemcpyHostToDevice, s->stream));
    k_retract<<<div_up(n, 128), 128, 0, s->stream>>>(t.ecm, t.bins, n, dxy.p, dok.p, dout.p, dedge.p);
    s->launches++;
    CUDA_TRY(s, cudaMemcpyAsync(out_ok, dok.p, n, cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(s, cudaMemcpyAsync(out_xy, dout.p, sizeof(float2) * n, cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(s, cudaMemcpyAsync(out_edge, dedge.p, sizeof(int) * n, cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(s, cudaStreamSynchronize(s->stream));
    dxy.free(); dout.free(); dedge.free(); dok.free();
    return ECMGPU_OK;
}

int ecmgpu_find_neighbors(ecmgpu_sim* s, int count, int* out_ids5, int* out_counts) {
    if (!s) return ECMGPU_ERR_INVALID;
    if (count < 0 || count > s->prm.max_agents || !out_ids5 || !out_counts) return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_find_neighbors: bad arguments");
    CUDA_TRY(s, cudaSetDevice(s->prm.device));
    if (s->n_slots > 0) {
        int rc = ensure_ready(s);
        if (rc) return rc;
        TickView t = make_view(s);
        if (s->strips_on) {
            if (s->local_transport && s->n_ranks > 1)
                return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_find_neighbors is not available with in-process strips");
            rc = enqueue_pack(s, t);
            if (rc) return rc;
            rc = enqueue_exchange(s, t);
            if (rc) return rc;
        }
        rc = enqueue_grid_build(s, t);
        if (rc) return rc;
        CUDA_TRY(s, cudaMemsetAsync(s->d_nbr.p, 0xff, sizeof(int) * 5 * (size_t)s->n_slots, s->stream));
        CUDA_TRY(s, cudaMemsetAsync(s->d_nbr_cnt.p, 0xff, sizeof(int) * (size_t)s->n_slots, s->stream));
        if (s->neighbor_mode == ECMGPU_NEIGHBORS_KDTREE) {
            // every active slot in ascending order with ONE shared, zero-filled output vector
            // (Simulator::FindNNearestNeighbors, Simulator.cpp:211-227, called the way ORCA::GetVelocity calls it)
            rc = enqueue_kd_build(s, t);
            if (rc) return rc;
            const KdQuery q = make_kd_query(s, false);
            CUDA_TRY(s, cudaMemsetAsync(q.cache, 0, sizeof(int) * 5, s->stream));
            k_kd_query<<<div_up(s->n_slots, 128), 128, 0, s->stream>>>(t, q, 0);
            k_kd_resolve<<<div_up(s->n_slots, 256), 256, 0, s->stream>>>(s->n_slots, s->d_active.p, q, s->d_nbr.p, s->d_nbr_cnt.p);
        } else {
            k_knn_query<<<div_up(s->n_slots + (s->strips_on ? 2 * s->cap_halo + s->cap_self : 0), 128), 128, 0, s->stream>>>(t);
            k_fallback<<<kSMs, 128, 0, s->stream>>>(t, 1);
        }
        s->launches += 2;
        CUDA_TRY(s, cudaGetLastError());
    }
    const int m = std::min(count, s->n_slots);
    if (m > 0) {  // the lists hold slot ids already (GridView::slot_of_row); the rows are indexed internally
        int rc = xfer(s, ECMGPU_NEIGHBORS, out_ids5, 0, m, true, false);
        if (rc) return rc;
        rc = xfer(s, ECMGPU_NEIGHBOR_COUNT, out_counts, 0, m, true, false);
        if (rc) return rc;
    }
    CUDA_TRY(s, cudaStreamSynchronize(s->stream));
    for (int i = m; i < count; i++) { out_counts[i] = -1; for (int j = 0; j < 5; j++) out_ids5[5 * i + j] = -1; }
    return ECMGPU_OK;
}

int ecmgpu_find_obstacles(ecmgpu_sim* s, int slot, int* out_ids, int cap, int* out_n) {
    if (!s) return ECMGPU_ERR_INVALID;
    if (slot < 0 || slot >= s->n_slots || cap < 0 || !out_n || (cap > 0 && !out_ids)) return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_find_obstacles: bad arguments");
    CUDA_TRY(s, cudaSetDevice(s->prm.device));
    int rc = ensure_ready(s);
    if (rc) return rc;
    float2 pos; float rad, spd;
    const int a = s->h_int_of[slot];
    CUDA_TRY(s, cudaMemcpyAsync(&pos, s->d_pos.p + a, sizeof(float2), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(s, cudaMemcpyAsync(&rad, s->d_radius.p + a, sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(s, cudaMemcpyAsync(&spd, s->d_speed.p + a, sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(s, cudaStreamSynchronize(s->stream));
    const float range = kLookAhead * spd + rad;
    DevBuf<int> dout, dn;
    CUDA_TRY(s, dout.alloc(std::max(cap, 1))); CUDA_TRY(s, dn.alloc(1));
    TickView t = make_view(s);
    k_find_obstacles<<<1, 32, 0, s->stream>>>(t.obst, t.bins, pos, range * range, dout.p, cap, dn.p);
    s->launches++;
    CUDA_TRY(s, cudaMemcpyAsync(out_n, dn.p, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(s, cudaStreamSynchronize(s->stream));
    const int m = std::min(*out_n, cap);
    if (m > 0) CUDA_TRY(s, cudaMemcpy(out_ids, dout.p, sizeof(int) * m, cudaMemcpyDeviceToHost));
    dout.free(); dn.free();
    return ECMGPU_OK;
}

int ecmgpu_set_ecm_topology(ecmgpu_sim* s, const int* vert_he, const int* he_next) {
    if (!s) return ECMGPU_ERR_INVALID;
    if (!s->have_ecm) return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_set_ecm_topology: call ecmgpu_set_ecm first");
    if (!vert_he || !he_next) return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_set_ecm_topology: null argument");
    const int nV = s->n_vertices, nH = 2 * s->n_edges;
    auto source = [&](int he) { return (he & 1) ? s->h_edge_v[2 * (he >> 1) + 1] : s->h_edge_v[2 * (he >> 1)]; };
    for (int v = 0; v < nV; v++)
        if (vert_he[v] < 0 || vert_he[v] >= nH || source(vert_he[v]) != v)
            return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_set_ecm_topology: vert_he[v] must be a half-edge leaving v");
    std::vector<unsigned char> seen(nH, 0);
    for (int h = 0; h < nH; h++) {
        if (he_next[h] < 0 || he_next[h] >= nH || source(he_next[h]) != source(h))
            return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_set_ecm_topology: he_next must stay on the half-edge's source vertex");
        // rings must close: the planner walks he_next until it is back where it started (AStar.cpp:104-152)
        if (seen[he_next[h]]++) return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_set_ecm_topology: he_next must be a permutation (closed rings)");
    }
    CUDA_TRY(s, cudaSetDevice(s->prm.device));
    CUDA_TRY(s, s->d_vert_he.alloc(nV)); CUDA_TRY(s, s->d_he_next.alloc(nH));
    CUDA_TRY(s, cudaMemcpyAsync(s->d_vert_he.p, vert_he, sizeof(int) * nV, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(s, cudaMemcpyAsync(s->d_he_next.p, he_next, sizeof(int) * nH, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(s, cudaStreamSynchronize(s->stream));
    s->have_topology = true;
    return ECMGPU_OK;
}

int ecmgpu_plan_paths(ecmgpu_sim* s, int n, const float* start_xy, const float* goal_xy, const float* clearance, int* out_off, int* out_len,
                      uint8_t* out_status, float* out_xy, int cap_points, int* out_points) {
    if (!s) return ECMGPU_ERR_INVALID;
    if (n < 0 || cap_points < 0 || (n > 0 && (!start_xy || !goal_xy || !clearance || !out_off || !out_len || (cap_points > 0 && !out_xy))))
        return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_plan_paths: bad arguments");
    if (!s->have_topology || !s->d_vert_clear.p)
        return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_plan_paths: needs vertex clearances (ecmgpu_set_ecm) and ecmgpu_set_ecm_topology");
    if (out_points) *out_points = 0;
    if (n == 0) return ECMGPU_OK;
    CUDA_TRY(s, cudaSetDevice(s->prm.device));
    int rc = ensure_bins(s);
    if (rc) return rc;
    rc = plan_alloc(s, s->pl, n, false);
    if (rc) return rc;
    const TickView t = make_view(s);
    PlanView w;
    w.ecm = t.ecm;
    w.bins = t.bins;
    w.vert_clear = s->d_vert_clear.p;
    w.vert_he = s->d_vert_he.p;
    w.he_next = s->d_he_next.p;
    DevBuf<float2> d_start, d_goal, d_pool;
    DevBuf<float> d_cl;
    DevBuf<int> d_off, d_len, d_cursor;
    DevBuf<unsigned char> d_st;
    CUDA_TRY(s, d_start.alloc(n)); CUDA_TRY(s, d_goal.alloc(n)); CUDA_TRY(s, d_cl.alloc(n)); CUDA_TRY(s, d_off.alloc(n));
    CUDA_TRY(s, d_len.alloc(n)); CUDA_TRY(s, d_st.alloc(n)); CUDA_TRY(s, d_cursor.alloc(1)); CUDA_TRY(s, d_pool.alloc(std::max(cap_points, 1)));
    CUDA_TRY(s, cudaMemcpyAsync(d_start.p, start_xy, sizeof(float2) * (size_t)n, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(s, cudaMemcpyAsync(d_goal.p, goal_xy, sizeof(float2) * (size_t)n, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(s, cudaMemcpyAsync(d_cl.p, clearance, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(s, cudaMemsetAsync(d_cursor.p, 0, sizeof(int), s->stream));
    const PlanScratch sc = make_plan_scratch(s->pl);
    if (!s->pl_ev[0]) { CUDA_TRY(s, cudaEventCreate(&s->pl_ev[0])); CUDA_TRY(s, cudaEventCreate(&s->pl_ev[1])); }
    CUDA_TRY(s, cudaEventRecord(s->pl_ev[0], s->stream));
    k_plan_paths<<<div_up(sc.n_workers, kPlanBlock), kPlanBlock, 0, s->stream>>>(w, sc, n, nullptr, d_start.p, d_goal.p, d_cl.p, d_off.p, d_len.p, d_st.p,
                                                                                 d_pool.p, cap_points, d_cursor.p);
    s->launches++;
    CUDA_TRY(s, cudaGetLastError());
    // second pass: the queries that ran out of first-pass capacity, with the full capacities (their first attempt
    // appended nothing to the pool)
    std::vector<unsigned char> st_host((size_t)n);
    CUDA_TRY(s, cudaMemcpyAsync(st_host.data(), d_st.p, (size_t)n, cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(s, cudaStreamSynchronize(s->stream));
    std::vector<int> redo;
    for (int q = 0; q < n; q++) if (st_host[q] == kPlanOverflow) redo.push_back(q);
    s->pl_last_workers = sc.n_workers;
    s->pl_last_second_pass = (int)redo.size();
    if (!redo.empty()) {
        ecmgpu_sim::PlanBufs full;
        DevBuf<int> d_redo;
        rc = plan_alloc(s, full, (int)redo.size(), true);
        if (!rc && d_redo.alloc(redo.size()) != cudaSuccess) rc = fail(s, ECMGPU_ERR_CUDA, "ecmgpu_plan_paths: out of device memory");
        if (!rc) {
            CUDA_TRY(s, cudaMemcpyAsync(d_redo.p, redo.data(), sizeof(int) * redo.size(), cudaMemcpyHostToDevice, s->stream));
            const PlanScratch sc2 = make_plan_scratch(full);
            k_plan_paths<<<div_up(sc2.n_workers, kPlanBlock), kPlanBlock, 0, s->stream>>>(w, sc2, (int)redo.size(), d_redo.p, d_start.p, d_goal.p, d_cl.p, d_off.p,
                                                                                           d_len.p, d_st.p, d_pool.p, cap_points, d_cursor.p);
            s->launches++;
            CUDA_TRY(s, cudaGetLastError());
            CUDA_TRY(s, cudaStreamSynchronize(s->stream));
        }
        full.free();
        d_redo.free();
        if (rc) return rc;
    }
    CUDA_TRY(s, cudaEventRecord(s->pl_ev[1], s->stream));
    int used = 0;
    CUDA_TRY(s, cudaMemcpyAsync(&used, d_cursor.p, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(s, cudaMemcpyAsync(out_off, d_off.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(s, cudaMemcpyAsync(out_len, d_len.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, s->stream));
    if (out_status) CUDA_TRY(s, cudaMemcpyAsync(out_status, d_st.p, (size_t)n, cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(s, cudaStreamSynchronize(s->stream));
    CUDA_TRY(s, cudaEventElapsedTime(&s->pl_last_ms, s->pl_ev[0], s->pl_ev[1]));
    const int have = std::min(used, cap_points);
    if (have > 0) CUDA_TRY(s, cudaMemcpy(out_xy, d_pool.p, sizeof(float2) * (size_t)have, cudaMemcpyDeviceToHost));
    if (out_points) *out_points = used;
    d_start.free(); d_goal.free(); d_pool.free(); d_cl.free(); d_off.free(); d_len.free(); d_cursor.free(); d_st.free();
    // a batch that took tens of gigabytes of scratch (the routes of a whole crowd) gives them back; the scratch of the
    // per-tick replans (a few hundred queries) stays for the next call
    const size_t held = (size_t)s->pl.workers * ((size_t)s->n_vertices * sizeof(PlanNode) + (size_t)s->pl.cap_push * 8 + (size_t)s->pl.cap_path * 8 +
                                                  (size_t)s->pl.cap_portals * 16 + (size_t)s->pl.cap_out * 8);
    if (held > ((size_t)2 << 30)) plan_free(s);
    return ECMGPU_OK;
}

int ecmgpu_plan_info(ecmgpu_sim* s, int* workers, float* kernel_ms, int* second_pass) {
    if (!s) return ECMGPU_ERR_INVALID;
    if (workers) *workers = s->pl_last_workers;
    if (kernel_ms) *kernel_ms = s->pl_last_ms;
    if (second_pass) *second_pass = s->pl_last_second_pass;
    return ECMGPU_OK;
}

int ecmgpu_valid_spawn_locations(ecmgpu_sim* s, int n, const float* xy, const float* clearance, uint8_t* out_valid) {
    if (!s) return ECMGPU_ERR_INVALID;
    if (n < 0 || (n > 0 && (!xy || !clearance || !out_valid))) return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_valid_spawn_locations: bad arguments");
    if (n == 0) return ECMGPU_OK;
    CUDA_TRY(s, cudaSetDevice(s->prm.device));
    if (s->n_slots == 0) {  // nobody to collide with (ensure_ready may have no crowd to size the grid from yet)
        memset(out_valid, 1, (size_t)n);
        return ECMGPU_OK;
    }
    if (s->strips_on && s->local_transport && s->n_ranks > 1)
        return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_valid_spawn_locations is not available with in-process strips");
    int rc = ensure_ready(s);
    if (rc) return rc;
    TickView t = make_view(s);
    if (s->strips_on) {
        rc = enqueue_pack(s, t);
        if (rc) return rc;
        rc = enqueue_exchange(s, t);
        if (rc) return rc;
    }
    rc = enqueue_grid_build(s, t);
    if (rc) return rc;
    DevBuf<float2> d_xy;
    DevBuf<float> d_c;
    DevBuf<unsigned char> d_out;
    CUDA_TRY(s, d_xy.alloc(n)); CUDA_TRY(s, d_c.alloc(n)); CUDA_TRY(s, d_out.alloc(n));
    CUDA_TRY(s, cudaMemcpyAsync(d_xy.p, xy, sizeof(float2) * (size_t)n, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(s, cudaMemcpyAsync(d_c.p, clearance, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, s->stream));
    k_valid_spawn<<<div_up(n, 128), 128, 0, s->stream>>>(t.grid, n, d_xy.p, d_c.p, d_out.p);
    s->launches++;
    CUDA_TRY(s, cudaGetLastError());
    CUDA_TRY(s, cudaMemcpyAsync(out_valid, d_out.p, (size_t)n, cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(s, cudaStreamSynchronize(s->stream));
    d_xy.free(); d_c.free(); d_out.free();
    return ECMGPU_OK;
}

int ecmgpu_draw_spawns(ecmgpu_sim* s, int n, const float* spawn_boxes, const float* goal_boxes, const float* clearance, uint64_t seed, uint64_t counter,
                       int max_attempts, float* out_start_xy, float* out_goal_xy, uint8_t* out_ok) {
    if (!s) return ECMGPU_ERR_INVALID;
    if (n < 0 || max_attempts < 1 || (n > 0 && (!spawn_boxes || !goal_boxes || !clearance || !out_start_xy || !out_goal_xy || !out_ok)))
        return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_draw_spawns: bad arguments");
    if (n == 0) return ECMGPU_OK;
    CUDA_TRY(s, cudaSetDevice(s->prm.device));
    if (s->strips_on && s->local_transport && s->n_ranks > 1)
        return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_draw_spawns is not available with in-process strips");
    TickView t = make_view(s);
    t.grid.n_sorted = 0;  // no crowd yet: every draw is valid
    if (s->n_slots > 0) {
        int rc = ensure_ready(s);
        if (rc) return rc;
        t = make_view(s);
        if (s->strips_on) {
            rc = enqueue_pack(s, t);
            if (rc) return rc;
            rc = enqueue_exchange(s, t);
            if (rc) return rc;
        }
        rc = enqueue_grid_build(s, t);
        if (rc) return rc;
        t.grid.n_sorted = 1;  // "there is a snapshot": the kernel only tests for zero
    }
    DevBuf<float4> d_sb, d_gb;
    DevBuf<float> d_c;
    DevBuf<float2> d_start, d_goal;
    DevBuf<unsigned char> d_ok;
    CUDA_TRY(s, d_sb.alloc(n)); CUDA_TRY(s, d_gb.alloc(n)); CUDA_TRY(s, d_c.alloc(n)); CUDA_TRY(s, d_start.alloc(n)); CUDA_TRY(s, d_goal.alloc(n)); CUDA_TRY(s, d_ok.alloc(n));
    CUDA_TRY(s, cudaMemcpyAsync(d_sb.p, spawn_boxes, sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(s, cudaMemcpyAsync(d_gb.p, goal_boxes, sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(s, cudaMemcpyAsync(d_c.p, clearance, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, s->stream));
    k_draw_spawns<<<div_up(n, 128), 128, 0, s->stream>>>(t.grid, n, d_sb.p, d_gb.p, d_c.p, seed, counter, max_attempts, d_start.p, d_goal.p, d_ok.p);
    s->launches++;
    CUDA_TRY(s, cudaGetLastError());
    CUDA_TRY(s, cudaMemcpyAsync(out_start_xy, d_start.p, sizeof(float2) * (size_t)n, cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(s, cudaMemcpyAsync(out_goal_xy, d_goal.p, sizeof(float2) * (size_t)n, cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(s, cudaMemcpyAsync(out_ok, d_ok.p, (size_t)n, cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(s, cudaStreamSynchronize(s->stream));
    return ECMGPU_OK;
}

// Back to "internal index = slot id": every per-agent array is permuted once.  The KD-tree mode needs it (it walks the
// slots in ascending order like KDTree::Construct and ORCA's carried neighbour list, KDTree.cpp:34-41, ORCA.h:100).
static int restore_identity(ecmgpu_sim* s) {
    if (s->perm_identity) return ECMGPU_OK;
    const int n = s->n_slots;
    if (n > 0) {
        DevBuf<unsigned char> tmp;
        CUDA_TRY(s, tmp.alloc(20 * (size_t)n));
        auto fix = [&](void* arr, size_t es) -> cudaError_t {
            const int nb = div_up(n, 256);
            const int* m = s->d_int_of.p;
            if (es == 1) k_gather_slots<unsigned char><<<nb, 256, 0, s->stream>>>(n, 0, m, (const unsigned char*)arr, (unsigned char*)tmp.p);
            else if (es == 4) k_gather_slots<unsigned><<<nb, 256, 0, s->stream>>>(n, 0, m, (const unsigned*)arr, (unsigned*)tmp.p);
            else if (es == 8) k_gather_slots<float2><<<nb, 256, 0, s->stream>>>(n, 0, m, (const float2*)arr, (float2*)tmp.p);
            else if (es == 16) k_gather_slots<float4><<<nb, 256, 0, s->stream>>>(n, 0, m, (const float4*)arr, (float4*)tmp.p);
            else k_gather_slots<Int5><<<nb, 256, 0, s->stream>>>(n, 0, m, (const Int5*)arr, (Int5*)tmp.p);
            s->launches++;
            return cudaMemcpyAsync(arr, tmp.p, es * (size_t)n, cudaMemcpyDeviceToDevice, s->stream);
        };
        static_assert(sizeof(PathHdr) == 16, "path headers move as float4");
        CUDA_TRY(s, fix(s->d_pos.p, 8)); CUDA_TRY(s, fix(s->d_vel.p, 8)); CUDA_TRY(s, fix(s->d_prefvel.p, 8));
        CUDA_TRY(s, fix(s->d_attraction.p, 8)); CUDA_TRY(s, fix(s->d_force.p, 8));
        CUDA_TRY(s, fix(s->d_radius.p, 4)); CUDA_TRY(s, fix(s->d_speed.p, 4)); CUDA_TRY(s, fix(s->d_status.p, 4));
        CUDA_TRY(s, fix(s->d_cell.p, 4)); CUDA_TRY(s, fix(s->d_nbr_cnt.p, 4));
        CUDA_TRY(s, fix(s->d_active.p, 1)); CUDA_TRY(s, fix(s->d_replan_pending.p, 1));
        CUDA_TRY(s, fix(s->d_nbr.p, 20)); CUDA_TRY(s, fix(s->d_path_hdr.p, 16));
        CUDA_TRY(s, cudaStreamSynchronize(s->stream));
        std::vector<PathHdr> hdr(s->h_path_hdr);
        for (int e = 0; e < n; e++) s->h_path_hdr[e] = hdr[s->h_int_of[e]];
    }
    for (size_t i = 0; i < s->h_int_of.size(); i++) s->h_int_of[i] = s->h_ext_of[i] = (int)i;
    CUDA_TRY(s, cudaMemcpyAsync(s->d_int_of.p, s->h_int_of.data(), sizeof(int) * s->h_int_of.size(), cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(s, cudaMemcpyAsync(s->d_ext_of.p, s->h_ext_of.data(), sizeof(int) * s->h_ext_of.size(), cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(s, cudaStreamSynchronize(s->stream));
    s->perm_identity = true;
    s->walk_dirty = true;
    s->config_epoch++;
    return ECMGPU_OK;
}

int ecmgpu_set_neighbor_mode(ecmgpu_sim* s, int mode) {
    if (!s) return ECMGPU_ERR_INVALID;
    if (mode != ECMGPU_NEIGHBORS_EXACT && mode != ECMGPU_NEIGHBORS_KDTREE) return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_set_neighbor_mode: unknown mode");
    CUDA_TRY(s, cudaSetDevice(s->prm.device));
    if (mode == ECMGPU_NEIGHBORS_KDTREE) {
        if (s->strips_on) return fail(s, ECMGPU_ERR_INVALID, "the KD-tree neighbour mode runs on a single handle (no strips)");
        int rc = restore_identity(s);
        if (rc) return rc;
        rc = kd_alloc(s);
        if (rc) return rc;
        CUDA_TRY(s, cudaMemsetAsync(s->d_kd_cache.p, 0, sizeof(int) * 10, s->stream));  // a fresh ORCA object (ORCA.h:87)
    }
    s->neighbor_mode = mode;
    s->config_epoch++;
    return ECMGPU_OK;
}

int ecmgpu_get_stats(ecmgpu_sim* s, ecmgpu_stats* o) {
    if (!s || !o) return ECMGPU_ERR_INVALID;
    CUDA_TRY(s, cudaSetDevice(s->prm.device));
    unsigned long long c[C_COUNT];
    CUDA_TRY(s, cudaMemcpyAsync(c, s->d_counters.p, sizeof(c), cudaMemcpyDeviceToHost, s->stream));
    int n_sorted = 0;
    if (!s->grid_dirty && s->d_cell_count.p && s->ticks > 0)
        CUDA_TRY(s, cudaMemcpyAsync(&n_sorted, s->d_cell_count.p + (size_t)s->gw * s->gh, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(s, cudaStreamSynchronize(s->stream));
    memset(o, 0, sizeof(*o));
    o->n_slots = s->n_slots;
    o->n_active = n_sorted;  // agents active at the start of the last tick
    o->grid_w = s->gw; o->grid_h = s->gh; o->neighbor_cell = s->cell;
    o->bins_w = s->bins_w; o->bins_h = s->bins_h; o->static_bin = s->static_bin;
    o->max_cell_list = s->max_cell_list; o->max_obstacle_list = s->max_obst_list;
    o->ticks = s->ticks; o->kernel_launches = s->launches;
    o->knn_fallbacks = c[C_TOTAL_FALLBACK]; o->obstacle_overflows = c[C_TOTAL_OBST_OVF]; o->lp3d_runs = c[C_TOTAL_LP3D];
    o->location_failures = c[C_TOTAL_LOCFAIL]; o->replans = c[C_TOTAL_REPLAN]; o->halo_misses = c[C_TOTAL_HALO_MISS];
    o->kd_median_ties = c[C_TOTAL_KD_TIES];
    o->kd_small_ties = c[C_TOTAL_KD_SMALL_TIES];
    o->event_overflows = c[C_TOTAL_EV_OVERFLOW];
    o->nonfinite_agent_ticks = c[C_TOTAL_NONFINITE];
    return ECMGPU_OK;
}

void ecmgpu_abi_sizes(int32_t out[4]) {
    out[0] = (int32_t)sizeof(ecmgpu_params);
    out[1] = (int32_t)sizeof(ecmgpu_stats);
    out[2] = (int32_t)sizeof(ecmgpu_agent_rec);
    out[3] = 22;  // members of ecmgpu_stats
    static_assert(sizeof(ecmgpu_stats) == 10 * 4 + 12 * 8, "ecmgpu_stats changed: update out[3] and the bindings");
}

int ecmgpu_set_profiling(ecmgpu_sim* s, int on) {
    if (!s) return ECMGPU_ERR_INVALID;
    if (s->profiling != (on != 0)) s->config_epoch++;  // the captured tick gains / loses its event-record nodes
    s->profiling = on != 0;
    s->ev_valid = false;
    return ECMGPU_OK;
}

int ecmgpu_last_tick_ms(ecmgpu_sim* s, float out_ms[4]) {
    if (!s || !out_ms) return ECMGPU_ERR_INVALID;
    if (!s->profiling || !s->ev_valid) return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_last_tick_ms: enable profiling and run a tick first");
    CUDA_TRY(s, cudaSetDevice(s->prm.device));
    CUDA_TRY(s, cudaEventSynchronize(s->ev[3]));
    CUDA_TRY(s, cudaEventElapsedTime(&out_ms[0], s->ev[0], s->ev[3]));
    CUDA_TRY(s, cudaEventElapsedTime(&out_ms[1], s->ev[0], s->ev[1]));
    CUDA_TRY(s, cudaEventElapsedTime(&out_ms[2], s->ev[1], s->ev[2]));
    CUDA_TRY(s, cudaEventElapsedTime(&out_ms[3], s->ev[2], s->ev[3]));
    return ECMGPU_OK;
}

int ecmgpu_last_tick_phases(ecmgpu_sim* s, float out_ms[5]) {
    if (!s || !out_ms) return ECMGPU_ERR_INVALID;
    float four[4];
    int rc = ecmgpu_last_tick_ms(s, four);
    if (rc) return rc;
    float fb = 0.0f;
    CUDA_TRY(s, cudaEventElapsedTime(&fb, s->ev[4], s->ev[3]));
    out_ms[0] = four[0]; out_ms[1] = four[1]; out_ms[2] = four[2]; out_ms[3] = four[3] - fb; out_ms[4] = fb;
    return ECMGPU_OK;
}

int ecmgpu_mark(ecmgpu_sim* s, int which) {
    if (!s) return ECMGPU_ERR_INVALID;
    if (which < 0 || which >= 8) return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_mark: mark index out of range");
    CUDA_TRY(s, cudaSetDevice(s->prm.device));
    CUDA_TRY(s, cudaEventRecord(s->marks[which], s->stream));
    return ECMGPU_OK;
}

int ecmgpu_mark_elapsed_ms(ecmgpu_sim* s, int a, int b, float* out_ms) {
    if (!s) return ECMGPU_ERR_INVALID;
    if (a < 0 || a >= 8 || b < 0 || b >= 8 || !out_ms) return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_mark_elapsed_ms: bad arguments");
    CUDA_TRY(s, cudaSetDevice(s->prm.device));
    CUDA_TRY(s, cudaEventSynchronize(s->marks[b]));
    CUDA_TRY(s, cudaEventElapsedTime(out_ms, s->marks[a], s->marks[b]));
    return ECMGPU_OK;
}

void* ecmgpu_stream(ecmgpu_sim* s) { return s ? (void*)s->stream : nullptr; }

int ecmgpu_comm_unique_id(uint8_t out_id[128]) {
    std::string err;
    if (!out_id) return ECMGPU_ERR_INVALID;
    if (!load_nccl(err)) return fail(nullptr, ECMGPU_ERR_COMM, err);
    UniqueId id;
    int r = g_nccl.GetUniqueId(&id);
    if (r != 0) return fail(nullptr, ECMGPU_ERR_COMM, std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r));
    memcpy(out_id, id.internal, 128);
    return ECMGPU_OK;
}

int ecmgpu_comm_init(ecmgpu_sim* s, const uint8_t nccl_unique_id[128], int rank, int n_ranks) {
    if (!s) return ECMGPU_ERR_INVALID;
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_comm_init: bad rank / n_ranks");
    CUDA_TRY(s, cudaSetDevice(s->prm.device));
    s->rank = rank;
    s->n_ranks = n_ranks;
    if (n_ranks > 1) {
        if (!nccl_unique_id) return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_comm_init: null unique id");
        std::string err;
        if (!load_nccl(err)) return fail(s, ECMGPU_ERR_COMM, err);
        UniqueId id;
        memcpy(id.internal, nccl_unique_id, 128);
        NCCL_TRY(s, g_nccl.CommInitRank(&s->nccl_comm, n_ranks, id, rank));
    }
    return ECMGPU_OK;
}

int ecmgpu_comm_init_local(ecmgpu_sim* s, int rank, int n_ranks, ecmgpu_sim* left, ecmgpu_sim* right) {
    if (!s) return ECMGPU_ERR_INVALID;
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_comm_init_local: bad rank / n_ranks");
    if ((rank > 0) != (left != nullptr) || (rank < n_ranks - 1) != (right != nullptr))
        return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_comm_init_local: neighbours must match the rank's position");
    CUDA_TRY(s, cudaSetDevice(s->prm.device));
    s->rank = rank;
    s->n_ranks = n_ranks;
    s->peer[0] = left;
    s->peer[1] = right;
    s->local_transport = true;
    if (!s->ev_packed) CUDA_TRY(s, cudaEventCreateWithFlags(&s->ev_packed, cudaEventDisableTiming));
    if (!s->ev_pulled) CUDA_TRY(s, cudaEventCreateWithFlags(&s->ev_pulled, cudaEventDisableTiming));
    CUDA_TRY(s, cudaEventRecord(s->ev_pulled, s->stream));
    return ECMGPU_OK;
}

int ecmgpu_comm_p2p_export(ecmgpu_sim* s, uint8_t out_handles[128]) {
    if (!s || !out_handles) return ECMGPU_ERR_INVALID;
    if (!s->strips_on || !s->d_recv[0].p) return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_comm_p2p_export: call ecmgpu_comm_set_strips first");
    CUDA_TRY(s, cudaSetDevice(s->prm.device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    for (int d = 0; d < 2; d++) {
        cudaIpcMemHandle_t h;
        CUDA_TRY(s, cudaIpcGetMemHandle(&h, s->d_recv[d].p));
        memcpy(out_handles + 64 * d, &h, 64);
    }
    return ECMGPU_OK;
}

int ecmgpu_comm_p2p_connect(ecmgpu_sim* s, const uint8_t* left_handles, const uint8_t* right_handles) {
    if (!s) return ECMGPU_ERR_INVALID;
    if (!s->strips_on) return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_comm_p2p_connect: call ecmgpu_comm_set_strips first");
    if ((s->rank > 0) != (left_handles != nullptr) || (s->rank < s->n_ranks - 1) != (right_handles != nullptr))
        return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_comm_p2p_connect: handles must match the rank's neighbours");
    CUDA_TRY(s, cudaSetDevice(s->prm.device));
    const uint8_t* src[2] = {left_handles ? left_handles + 64 : nullptr,   // left neighbour's RIGHT inbox
                             right_handles ? right_handles : nullptr};     // right neighbour's LEFT inbox
    for (int d = 0; d < 2; d++) {
        if (!src[d]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, src[d], 64);
        void* p = nullptr;
        CUDA_TRY(s, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        s->peer_inbox[d] = (unsigned char*)p;
    }
    CUDA_TRY(s, s->d_send_hdr.alloc(2));
    CUDA_TRY(s, cudaMemsetAsync(s->d_send_hdr.p, 0, 2 * sizeof(MsgHeader), s->stream));
    CUDA_TRY(s, s->d_seq.alloc(1));
    const int seq = (int)s->comm_seq;
    CUDA_TRY(s, cudaMemcpyAsync(s->d_seq.p, &seq, sizeof(int), cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(s, cudaStreamSynchronize(s->stream));
    s->p2p = true;
    s->config_epoch++;
    return ECMGPU_OK;
}

int ecmgpu_comm_set_strips(ecmgpu_sim* s, const float* bounds, float halo_width) {
    if (!s) return ECMGPU_ERR_INVALID;
    if (!bounds || !(halo_width > 0.0f)) return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_comm_set_strips: bad arguments");
    if (s->neighbor_mode != ECMGPU_NEIGHBORS_EXACT) return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_comm_set_strips: the KD-tree neighbour mode runs on a single handle (no strips)");
    if (s->n_ranks > 1 && !s->nccl_comm && !s->local_transport) return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_comm_set_strips: call ecmgpu_comm_init first");
    for (int r = 0; r < s->n_ranks; r++) {
        if (!(bounds[r] < bounds[r + 1])) return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_comm_set_strips: bounds must ascend");
        // ghosts come from the adjacent strips only: an interior strip narrower than the halo would hide agents two strips away
        if (r > 0 && r < s->n_ranks - 1 && bounds[r + 1] - bounds[r] < halo_width)
            return fail(s, ECMGPU_ERR_INVALID, "ecmgpu_comm_set_strips: interior strip narrower than the halo");
    }
    CUDA_TRY(s, cudaSetDevice(s->prm.device));
    const float inf = std::numeric_limits<float>::infinity();
    s->strip_lo = s->rank == 0 ? -inf : bounds[s->rank];
    s->strip_hi = s->rank == s->n_ranks - 1 ? inf : bounds[s->rank + 1];
    s->halo = halo_width;
    // capacities of the fixed-size messages (counts travel in the header, no host round trip): four times
    // what the current crowd puts within the halo of ANY vertical line, never less than 4096 entries.  Taking the
    // worst line instead of today's borders makes the layout the same on every rank (every rank still sees
    // the whole crowd here) and keeps it valid when the borders move (re-balancing) or the crowd drifts.
    const int n = s->prm.max_agents;
    int near = 0;
    if (s->n_slots > 0) {
        std::vector<float2> pos(s->n_slots);
        std::vector<unsigned char> act(s->n_slots);
        CUDA_TRY(s, cudaMemcpyAsync(pos.data(), s->d_pos.p, sizeof(float2) * s->n_slots, cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(s, cudaMemcpyAsync(act.data(), s->d_active.p, s->n_slots, cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(s, cudaStreamSynchronize(s->stream));
        std::vector<float> xs;
        xs.reserve(s->n_slots);
        for (int i = 0; i < s->n_slots; i++)
            if (act[i]) xs.push_back(pos[i].x);
        std::sort(xs.begin(), xs.end());
        size_t j = 0;
        for (size_t i = 0; i < xs.size(); i++) {  // agents with x in [xs[i] - 2 halo, xs[i]]: within the halo of the line in the middle
            while (xs[i] - xs[j] > 2.0f * halo_width) j++;
            near = std::max(near, (int)(i - j + 1));
        }
    }
    const int want_halo = std::min(n, std::max(4096, 4 * near));
    if (s->strips_on) {
        // Re-balance: new borders for a crowd whose global state the caller has just written to every rank's slot
        // arrays.  The message buffers stay (peer-transport mappings and sequence numbers with them), so the
        // capacities chosen at the first call must still do.
        if (want_halo > s->cap_halo)
            return fail(s, ECMGPU_ERR_CAPACITY, "ecmgpu_comm_set_strips: the crowd near the new borders needs larger messages than "
                        "the first call allocated; create the strips on a fresh handle");
    } else {
        s->cap_halo = want_halo;
        s->cap_migr = std::min(n, std::max(1024, s->cap_halo / 4));
        s->cap_self = 2 * s->cap_migr;
        const size_t msg = strip_msg_bytes(s->cap_halo, s->cap_migr);
        for (int d = 0; d < 2; d++) {
            CUDA_TRY(s, s->d_send[d].alloc(msg));
            CUDA_TRY(s, s->d_recv[d].alloc(2 * msg));  // two generations for the peer transport
            CUDA_TRY(s, cudaMemsetAsync(s->d_send[d].p, 0, sizeof(MsgHeader), s->stream));
            CUDA_TRY(s, cudaMemsetAsync(s->d_recv[d].p, 0, 2 * msg, s->stream));
        }
        CUDA_TRY(s, s->d_self_ghost.alloc(s->cap_self));
        CUDA_TRY(s, s->d_self_ghost_n.alloc(1));
        CUDA_TRY(s, s->d_g_key.alloc(2 * (size_t)s->cap_halo + s->cap_self));
        CUDA_TRY(s, s->d_g_rank.alloc(2 * (size_t)s->cap_halo + s->cap_self));
        // snapshot arrays must hold owned agents + ghosts
        const size_t cap = (size_t)n + 2 * (size_t)s->cap_halo + s->cap_self;
        CUDA_TRY(s, s->d_s_pos.alloc(cap)); CUDA_TRY(s, s->d_s_vel.alloc(cap)); CUDA_TRY(s, s->d_s_pref.alloc(cap));
        CUDA_TRY(s, s->d_s_rad.alloc(cap)); CUDA_TRY(s, s->d_s_spd.alloc(cap)); CUDA_TRY(s, s->d_s_slot.alloc(cap));
        CUDA_TRY(s, s->d_s_alive.alloc(cap)); CUDA_TRY(s, s->d_s_ghost.alloc(cap)); CUDA_TRY(s, s->d_fb_list.alloc(cap));
    }
    s->strips_on = true;
    s->io.owned_confirmed = -1;
    s->walk_dirty = true;  // ownership is re-derived below
    if (s->compact) s->grid_dirty = true;  // the rank's grid follows its strip (build_grid)
    s->config_epoch++;
    if (s->n_slots > 0) {
        TickView t = make_view(s);
        k_assign_owner<<<div_up(s->n_slots, 256), 256, 0, s->stream>>>(s->n_slots, t.ag, make_strip_view(s));
        s->launches++;
        CUDA_TRY(s, cudaGetLastError());
    }
    CUDA_TRY(s, cudaStreamSynchronize(s->stream));
    return ECMGPU_OK;
}

}  // extern "C"
