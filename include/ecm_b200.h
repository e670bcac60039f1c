/* ecm_b200.h - C ABI of the B200 (sm_100a) implementation of ECMGenerator's per-tick agent update.
 *
 * The reference has no plugin/FFI mechanism; the seam this library sits behind is the C++ class
 * ECM::Simulation::Simulator (/root/reference/ECMAgentSimulator/Simulator.h:59-188).  A host-side
 * drop-in `Simulator` (ecmgenerator_b200/csrc/dropin/Simulator.h) keeps that class's public
 * signatures and forwards to the entry points below; INTEGRATION.md shows the binding.
 *
 * Conventions: plain C types, caller-owned host buffers, int status codes (0 = ECMGPU_OK),
 * one handle = one logical simulator on one CUDA device, calls on a handle serialised by the
 * caller.  Work is enqueued on the handle's CUDA stream; ecmgpu_update() returns without waiting,
 * ecmgpu_read()/ecmgpu_poll_events()/ecmgpu_sync() wait.  There is NO CPU fallback: without a
 * usable CUDA device every compute entry point fails with ECMGPU_ERR_CUDA.
 *
 * Each entry point cites the reference interface it replaces (paths relative to /root/reference).
 */
#ifndef ECM_B200_H
#define ECM_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ecmgpu_sim ecmgpu_sim;

enum {
    ECMGPU_OK = 0,
    ECMGPU_ERR_INVALID = 1,   /* bad argument / call order */
    ECMGPU_ERR_CUDA = 2,      /* CUDA runtime error or no device */
    ECMGPU_ERR_CAPACITY = 3,  /* slot, path pool or event capacity exceeded */
    ECMGPU_ERR_COMM = 4       /* NCCL error */
};

/* Hard constants of the reference are NOT parameters: k = 5 neighbours (Simulator.cpp:55),
 * look-ahead 10 s (ORCA.h:102-103), mass 0.8 (Simulator.cpp:622), arrival radius 20 / delete
 * distance 2 (Simulator.cpp:542-543), EPSILON 1e-4 (Configuration.h:14). */
typedef struct ecmgpu_params {
    int   device;            /* CUDA device ordinal */
    int   max_agents;        /* Simulator ctor `maxAgents`   (Simulator.h:63) */
    float step;              /* Simulator ctor `simStepTime` (Simulator.h:63); Update(dt) ignores dt (Simulator.cpp:314) */
    float neighbor_cell;     /* neighbour-grid cell edge in world units; <= 0: chosen from crowd density at load */
    float static_bin;        /* point-location / obstacle bin edge; <= 0: chosen from the ECM */
    float max_obstacle_range;/* upper bound of 10*speed + radius over all agents (ORCA.cpp:27); <= 0: tracked from loaded agents */
    int   path_pool_points;  /* capacity of the path polyline pool in points; <= 0: 16 * max_agents */
    int   record_neighbors;  /* 1: keep each tick's neighbour ids/counts readable (parity); 0: skip the 24 B/agent write */
} ecmgpu_params;

/* which-array selectors for ecmgpu_read / ecmgpu_write.  Element layout per slot in brackets. */
enum {
    ECMGPU_POS = 0,          /* [2 f32] Simulator::GetPositionData           (Simulator.h:116) */
    ECMGPU_VEL = 1,          /* [2 f32] Simulator::GetVelocityData           (Simulator.h:117) */
    ECMGPU_PREFVEL = 2,      /* [2 f32] Simulator::GetPreferredVelocityData  (Simulator.h:118) */
    ECMGPU_ATTRACTION = 3,   /* [2 f32] Simulator::GetAttractionPointData    (Simulator.h:121) */
    ECMGPU_FORCE = 4,        /* [2 f32] m_Forces (no public getter)          (Simulator.h:184) */
    ECMGPU_RADIUS = 5,       /* [1 f32] Simulator::GetClearanceData          (Simulator.h:120) */
    ECMGPU_SPEED = 6,        /* [1 f32] m_PreferredSpeed                     (Simulator.h:186) */
    ECMGPU_ACTIVE = 7,       /* [1 u8 ] Simulator::GetActiveFlags            (Simulator.h:123) */
    ECMGPU_CELL = 8,         /* [1 i32] ECM cell located this tick (2*edge+side), -1 none, -2 not evaluated (ECM.cpp:220-223) */
    ECMGPU_NEIGHBORS = 9,    /* [5 i32] neighbour slot ids of this tick, (sqDist, slot) ascending, -1 padding (KDTree.cpp:85-96) */
    ECMGPU_NEIGHBOR_COUNT = 10, /* [1 i32] */
    ECMGPU_STATUS = 11,      /* [1 u32] ECMGPU_ST_* bits of the last tick */
    ECMGPU_REPLAN_PENDING = 12 /* [1 u8]  a replan event was raised and ecmgpu_set_path has not answered it yet
                                *         (state a strip re-balance has to carry along with the agent) */
};

/* per-agent status bits (ECMGPU_STATUS) */
enum {
    ECMGPU_ST_NO_CELL = 1u,        /* "Couldn't locate point in the ECM graph!" (ECMCellCollection.cpp:88) */
    ECMGPU_ST_REPLAN = 2u,         /* IRM failed -> UpdatePath requested      (Simulator.cpp:581-587) */
    ECMGPU_ST_ARRIVING = 4u,       /* within the arrival radius                (Simulator.cpp:557-562) */
    ECMGPU_ST_DESTROYED = 8u,      /* destroyed this tick                      (Simulator.cpp:564-566) */
    ECMGPU_ST_OBST_OVERFLOW = 16u, /* more obstacle neighbours than the device cap (excess dropped) */
    ECMGPU_ST_KNN_FALLBACK = 32u,  /* neighbour search left the ring budget; resolved by the exhaustive pass */
    ECMGPU_ST_LP3D = 64u,          /* RandomizedLP failed, RandomizedLP3D ran  (ORCA.cpp:51-54) */
    ECMGPU_ST_HALO_MISS = 128u,    /* multi-GPU: search reached beyond the received halo */
    ECMGPU_ST_NONFINITE = 256u     /* the agent's position is NaN or infinite (the reference's LP3D can produce NaN velocities;
                                    * from there the reference is undefined): the agent stays active, is nobody's neighbour
                                    * and is no longer updated */
};

/* -- lifetime ---------------------------------------------------------------------------------
 * Simulator::Simulator + Initialize (Simulator.h:63-74, Simulator.cpp:21-60) / ~Simulator. */
int  ecmgpu_create(const ecmgpu_params* params, ecmgpu_sim** out);
void ecmgpu_destroy(ecmgpu_sim* sim);
/* Last error text of this handle (or of the failed ecmgpu_create when sim == NULL). */
const char* ecmgpu_last_error(const ecmgpu_sim* sim);

/* -- static world -----------------------------------------------------------------------------
 * Flattened ECMGraph (ECM.h:20-116): vertices, edges v0->v1 with the four closest obstacle points
 * L0 R0 L1 R1 (edge_cl: 8 floats per edge; layout in csrc/host/flat_world.h).  Cell c = 2*edge+side
 * as built by ECMCellCollection::Construct (ECMCellCollection.cpp:10-49).  bbox = walkable area. */
int ecmgpu_set_ecm(ecmgpu_sim* sim, const float bbox[4], int n_vertices, const float* vert_xy, const float* vert_clear,
                   int n_edges, const int* edge_v, const float* edge_cl);
/* Environment::GetObstacles() flattened in (obstacle, vertex) order (Environment.h:55,
 * ECMDataTypes.h:160-166): segment i runs xy[i] -> xy[next[i]]. */
int ecmgpu_set_obstacles(ecmgpu_sim* sim, int n, const float* xy, const int* next, const int* prev, const uint8_t* convex);

/* -- agents -----------------------------------------------------------------------------------
 * The host owns slot allocation (free list, Simulator.h:66-69); these write one slot's components.
 * Simulator::SpawnAgent after ValidSpawnLocation and path planning (Simulator.cpp:175-199). */
int ecmgpu_spawn(ecmgpu_sim* sim, int slot, float x, float y, float radius, float speed, const float* path_xy, int n_points);
/* n agents at once; slots == NULL means slots 0..n-1.  path_off[n+1] indexes points in path_xy. */
int ecmgpu_bulk_load(ecmgpu_sim* sim, int n, const int* slots, const float* pos_xy, const float* radius, const float* speed,
                     const int* path_off, const float* path_xy);
/* Simulator::UpdatePath's result (Simulator.cpp:97-124): replaces the polyline, clears a pending replan. */
int ecmgpu_set_path(ecmgpu_sim* sim, int slot, const float* path_xy, int n_points);
/* Simulator::DestroyAgent (Simulator.cpp:202-208): clears the active flag. */
int ecmgpu_destroy_agent(ecmgpu_sim* sim, int slot);

/* -- the hot path -----------------------------------------------------------------------------
 * Simulator::Update (Simulator.cpp:314-323) minus UpdateSpawnAreas (host): neighbour structure,
 * attraction points, preferred velocities, ORCA, velocity and position integration.  Asynchronous. */
int ecmgpu_update(ecmgpu_sim* sim);
int ecmgpu_sync(ecmgpu_sim* sim);
/* Drains the device event queues (waits for enqueued ticks): agents that asked for a replan
 * (Simulator.cpp:581-587; they stay flagged until ecmgpu_set_path) and agents destroyed on arrival
 * (Simulator.cpp:564-566) since the last poll.  Either array may be NULL with cap 0 to only count.
 * Each list holds max_agents entries: a host that reuses slots (respawn) without polling for more than max_agents
 * events loses the excess (counted in ecmgpu_stats.event_overflows) and the next poll fails with ECMGPU_ERR_CAPACITY. */
int ecmgpu_poll_events(ecmgpu_sim* sim, int* replan_slots, int replan_cap, int* n_replans, int* destroyed_slots,
                       int destroyed_cap, int* n_destroyed);

/* -- state transfer ---------------------------------------------------------------------------
 * Component arrays indexed by slot, [first, first+count).  read waits for enqueued work. */
int ecmgpu_read(ecmgpu_sim* sim, int which, void* dst, int first, int count);
int ecmgpu_write(ecmgpu_sim* sim, int which, const void* src, int first, int count);
/* Asynchronous variants on the handle's stream for pinned host memory (end-to-end pipelines). */
int ecmgpu_read_async(ecmgpu_sim* sim, int which, void* dst_pinned, int first, int count);
int ecmgpu_write_async(ecmgpu_sim* sim, int which, const void* src_pinned, int first, int count);
/* One tick with host I/O, pipelined over two copy streams and double-buffered device staging:
 * uploads positions / velocities of slots [0,count) from PINNED host memory (NULL = keep the device
 * values), runs ecmgpu_update, downloads positions / velocities / active flags into PINNED host
 * memory (NULL = skip).  Returns at once; the outputs of the call that returned `ticket` are complete
 * after ecmgpu_io_wait(ticket).  Consecutive calls overlap (upload of k+1 | tick k | download of k-1):
 * give every call in flight its own set of host buffers.  Two sets (wait for ticket k-1 after issuing
 * call k) already hide the tick behind the copies; three sets (wait for k-2) also let the download of
 * k-1 and the upload of k+1 share the link in both directions.  The last 8 tickets can be waited for
 * individually; an older ticket waits for the newest call that reused its slot.
 * Replaces the per-frame pattern Update() + GetPositionData()/GetVelocityData()/GetActiveFlags()
 * (Application.cpp:137-144, ECMRenderer.cpp:836-884) for hosts that keep the state on their side. */
int ecmgpu_update_io(ecmgpu_sim* sim, int count, const float* in_pos, const float* in_vel, float* out_pos, float* out_vel,
                     uint8_t* out_active, uint64_t* ticket);
int ecmgpu_io_wait(ecmgpu_sim* sim, uint64_t ticket);
/* The same pipeline for strips (and for hosts that only track live agents): moves (slot, position,
 * velocity) records of the agents THIS handle owns instead of whole slot arrays, so that with N ranks
 * every rank transfers its share of the crowd.  `in` (PINNED, may be NULL with n_in = 0): records whose
 * slot this handle owns overwrite position and velocity before the tick, other records are ignored.
 * After the tick `out` (PINNED, room for out_cap records) receives one record per owned agent, in no
 * particular order, and *out_count (PINNED) their number.  The records are staged on the device and the
 * copy is sized from the count confirmed by the last ecmgpu_io_wait plus the migrants that may have arrived
 * since (the first calls after a spawn / bulk load move max_agents records).  ECMGPU_IO_DIRECT=1 (opt-in,
 * slower over PCIe) stores them straight into `out` when that is pinned memory the device can address.
 * ecmgpu_io_wait fails with ECMGPU_ERR_CAPACITY when out_cap was too small for the tick's count (nothing is
 * written past out_cap). */
typedef struct ecmgpu_agent_rec {
    int32_t slot;
    float x, y, vx, vy;
} ecmgpu_agent_rec; /* 20 bytes */
int ecmgpu_update_io_owned(ecmgpu_sim* sim, int n_in, const ecmgpu_agent_rec* in, ecmgpu_agent_rec* out, int out_cap,
                           int32_t* out_count, uint64_t* ticket);
/* cudaHostAlloc / cudaFreeHost pass-through so non-CUDA hosts can get pinned staging memory. */
void* ecmgpu_alloc_pinned(uint64_t bytes);
void  ecmgpu_free_pinned(void* p);

/* -- queries (Simulator.h:106-108, ECM.h:127-128) on the CURRENT state ---------------------------
 * ECM::GetECMCell for arbitrary points. */
int ecmgpu_locate(ecmgpu_sim* sim, int n, const float* xy, int* out_cell);
/* ECM::RetractPoint for arbitrary points. */
int ecmgpu_retract(ecmgpu_sim* sim, int n, const float* xy, uint8_t* out_ok, float* out_xy, int* out_edge);
/* Simulator::FindNNearestNeighbors(k=5) for every active slot < count on the current positions
 * (exact-kNN contract); inactive slots get count -1. */
int ecmgpu_find_neighbors(ecmgpu_sim* sim, int count, int* out_ids5, int* out_counts);
/* Simulator::FindNearestObstacles with ORCA's range for one slot; returns the number found. */
int ecmgpu_find_obstacles(ecmgpu_sim* sim, int slot, int* out_ids, int cap, int* out_n);

/* Simulator::ValidSpawnLocation (Simulator.cpp:295-311) for n candidate points at once, on the CURRENT positions:
 * out_valid[i] = 1 iff no active agent centre lies strictly closer to xy[i] than clearance[i] (the reference's float
 * expression).  Uses the neighbour grid instead of the reference's scan over every agent; candidates of one batch do
 * not see each other (a host that spawns several agents per tick checks those few against each other itself, as
 * UpdateSpawnAreas does one by one, Simulator.cpp:494-536).  With strips: against the agents this handle holds
 * (owned + halo), and - like ecmgpu_find_neighbors - a collective call: every rank runs the halo exchange. */
int ecmgpu_valid_spawn_locations(ecmgpu_sim* sim, int n, const float* xy, const float* clearance, uint8_t* out_valid);

/* Simulator::UpdateSpawnAreas' draws (Simulator.cpp:501-527, Area::GetRandomPositionInArea Area.h:39-53) on the device with
 * a counter-based generator instead of C rand() - opt-in, for hosts that do not need the reference's rand() stream: request i
 * tries up to max_attempts positions uniform in spawn_boxes[4i..4i+3] (xmin ymin xmax ymax) and keeps the first that passes
 * Simulator::ValidSpawnLocation on the CURRENT positions (out_ok[i] = 0: none did, out_start is the last try); its goal is
 * uniform in goal_boxes[4i..]. Every draw is a pure function of (seed, counter, i, attempt): reproducible, independent of
 * batching, NOT the reference's stream (parity is statistical).  Requests of one call do not see each other; the caller
 * resolves those few conflicts in request order, as UpdateSpawnAreas does one by one.  Collective with strips. */
int ecmgpu_draw_spawns(ecmgpu_sim* sim, int n, const float* spawn_boxes, const float* goal_boxes, const float* clearance, uint64_t seed,
                       uint64_t counter, int max_attempts, float* out_start_xy, float* out_goal_xy, uint8_t* out_ok);

/* -- batched path planning on the device -------------------------------------
 * The half-edge rings of the ECM graph, needed by the planner only: vert_he[v] = one half-edge leaving vertex v
 * (ECMVertex::half_edge_idx, ECM.h:41-47), he_next[h] = the next half-edge around h's source vertex
 * (ECMHalfEdge::next_idx, ECM.h:20-26); half-edge 2e runs edge e from v0 to v1, 2e+1 back (ECM.h:69).
 * ecmgpu_set_ecm must have been given the vertex clearances.  A new ecmgpu_set_ecm drops the topology. */
int ecmgpu_set_ecm_topology(ecmgpu_sim* sim, const int* vert_he, const int* he_next);
/* ECMPathPlanner::FindPath (ECMPathPlanner.cpp:22-136, preferredAdditionalClearance = 0 as in Simulator.cpp:108-112)
 * for n queries at once, one device thread per query: point location, retraction, A* on the medial axis with the
 * reference's open-list behaviour, corridor, portals, funnel - the polylines equal the reference's bit for bit
 * (tests/test_hostdev_planner.py).  The polyline of query i is out_xy[2*out_off[i] .. 2*(out_off[i]+out_len[i]));
 * paths are packed in order of completion, so out_off is NOT ascending in i.  out_len[i] = 0 where the reference
 * returns false.  out_status (may be NULL): 0 path, 1 no path, 2 a capacity was exceeded (path longer than 1024
 * points / 2048 graph vertices / 8192 portals, or the pool: *out_points > cap_points - call again with a larger
 * pool).  Waits for the result. */
int ecmgpu_plan_paths(ecmgpu_sim* sim, int n, const float* start_xy, const float* goal_xy, const float* clearance, int* out_off,
                      int* out_len, uint8_t* out_status, float* out_xy, int cap_points, int* out_points);
/* Measurement aid for the call above: how many queries the last ecmgpu_plan_paths kept in flight (workers with their
 * own scratch), the device time of its kernels in milliseconds (CUDA events on the simulator's stream; the copies of
 * the queries and of the polylines are not in it) and how many queries needed the second pass (the first one runs with
 * scratch sized for the usual query; a query that exceeds it is planned again with the full capacities).  Any pointer
 * may be NULL. */
int ecmgpu_plan_info(ecmgpu_sim* sim, int* workers, float* kernel_ms, int* second_pass);

/* -- neighbour mode ---------------------------------------------------------------------------
 * ECMGPU_NEIGHBORS_EXACT (default): the exact 5-NN contract of DESIGN.md on the per-tick uniform grid.
 * ECMGPU_NEIGHBORS_KDTREE: the reference's own lists - a median-split tree built like KDTree::Construct
 * (KDTree.cpp:22-83) and searched like KDTree::KNearestAgents_R (KDTree.cpp:98-202) with its pruning and
 * fill rules and the ids ORCA::m_NeighborCache (ORCA.h:100) carries from one query to the next, so that
 * trajectories can be compared with the UNMODIFIED reference.  A parity mode: single handle (no strips),
 * several times the cost of the default.  Switching modes resets the carried-over list to zeros, the
 * state of a fresh ORCA object (ORCA.h:87).  ECMGPU_NEIGHBORS then holds the tick's lists in the
 * reference's order (place 0 = farthest), ecmgpu_find_neighbors answers in the same mode. */
enum { ECMGPU_NEIGHBORS_EXACT = 0, ECMGPU_NEIGHBORS_KDTREE = 1 };
int ecmgpu_set_neighbor_mode(ecmgpu_sim* sim, int mode);

/* -- introspection ----------------------------------------------------------------------------*/
typedef struct ecmgpu_stats {
    int   n_slots;            /* highest loaded slot + 1 (what the kernels iterate over) */
    int   n_active;           /* active agents after the last completed tick */
    int   grid_w, grid_h;     /* neighbour grid */
    float neighbor_cell;
    int   bins_w, bins_h;     /* static grid */
    float static_bin;
    int   max_cell_list, max_obstacle_list;
    uint64_t ticks;           /* ticks enqueued so far */
    uint64_t kernel_launches; /* kernels launched by this handle so far */
    uint64_t knn_fallbacks, obstacle_overflows, lp3d_runs, location_failures, replans, halo_misses;
    /* ECMGPU_NEIGHBORS_KDTREE: tree segments whose median tied with another agent on the split axis.  In segments of
     * up to 16 agents (kd_small_ties) the tie resolves like libstdc++'s std::sort does (insertion sort, stable);
     * in larger ones (kd_median_ties) the reference's tree is std::sort-defined: 0 = the tree equals the reference's */
    uint64_t kd_median_ties, kd_small_ties;
    /* arrival / replan events dropped because a list was full: each list holds max_agents entries and is emptied by
     * ecmgpu_poll_events, which then fails with ECMGPU_ERR_CAPACITY (poll at least once per max_agents events) */
    uint64_t event_overflows;
    /* agent-ticks skipped because the agent's position was not finite (ECMGPU_ST_NONFINITE) */
    uint64_t nonfinite_agent_ticks;
} ecmgpu_stats;
int ecmgpu_get_stats(ecmgpu_sim* sim, ecmgpu_stats* out);
/* sizeof(ecmgpu_params), sizeof(ecmgpu_stats), sizeof(ecmgpu_agent_rec), number of ecmgpu_stats members: lets a binding in
 * another language check its mirrors of the structs against THIS build (needs no device). */
void ecmgpu_abi_sizes(int32_t out[4]);
/* CUDA-event time of each phase of the LAST completed tick, milliseconds.
 * phases: 0 whole tick, 1 grid build (count+scan+scatter), 2 attraction (locate+IRM+steer), 3 ORCA (kNN+obstacles+LP+integrate) */
int ecmgpu_last_tick_ms(ecmgpu_sim* sim, float out_ms[4]);
/* The same with phase 3 split: 0 whole tick, 1 grid build, 2 attraction, 3 k_orca alone (default neighbour mode; in the
 * KD-tree mode: tree build + search + ORCA), 4 k_fallback (stragglers of the neighbour search and RandomizedLP3D of the
 * agents k_orca parked: 13 % of a congested crowd). */
int ecmgpu_last_tick_phases(ecmgpu_sim* sim, float out_ms[5]);
/* Enables per-phase event recording: 5 event records per tick, inside the captured graph of the tick (external
 * event-record nodes), so the phases are those of the tick an unprofiled run replays. */
int ecmgpu_set_profiling(ecmgpu_sim* sim, int on);
/* Stream-ordered time marks (CUDA events on the handle's stream): record mark `which` (0..7) now;
 * elapsed waits for mark b and returns the device time between marks a and b in milliseconds. */
int ecmgpu_mark(ecmgpu_sim* sim, int which);
int ecmgpu_mark_elapsed_ms(ecmgpu_sim* sim, int a, int b, float* out_ms);
/* The stream work is enqueued on (a cudaStream_t), for callers that interleave their own CUDA work. */
void* ecmgpu_stream(ecmgpu_sim* sim);

/* -- multi-GPU strips (one process per GPU; see DESIGN.md "Multi-GPU") --------------------------
 * nccl_unique_id: 128 bytes from ecmgpu_comm_unique_id() on rank 0, distributed by the caller. */
int ecmgpu_comm_unique_id(uint8_t out_id[128]);
int ecmgpu_comm_init(ecmgpu_sim* sim, const uint8_t nccl_unique_id[128], int rank, int n_ranks);
/* Peer transport (one process per GPU, same node): after ecmgpu_comm_set_strips every rank exports the
 * CUDA-IPC handles of its two inboxes, the caller hands each rank its neighbours' 128 bytes, and from
 * then on k_pack stores halo / migrant entries straight into the neighbour's inbox over NVLink; a
 * sequence number written after a system-scope fence replaces the NCCL send/recv pair.
 * left/right = the neighbour's exported bytes, NULL at the rim. */
int ecmgpu_comm_p2p_export(ecmgpu_sim* sim, uint8_t out_handles[128]);
int ecmgpu_comm_p2p_connect(ecmgpu_sim* sim, const uint8_t* left_handles, const uint8_t* right_handles);
/* In-process transport instead of NCCL: all strips are handles of ONE process (on one or several
 * devices); messages move by peer copies.  left/right are the neighbouring handles (NULL at the rim).
 * With this transport a tick is driven as: ecmgpu_update_phase(h, 0) on every handle, then phase 1
 * on every handle, then phase 2 on every handle. */
int ecmgpu_comm_init_local(ecmgpu_sim* sim, int rank, int n_ranks, ecmgpu_sim* left, ecmgpu_sim* right);
/* One tick in three stream-ordered phases: 0 = pack halo / migrants, 1 = exchange + adopt migrants,
 * 2 = neighbour grid, attraction, ORCA, integration.  ecmgpu_update() = phases 0, 1, 2. */
int ecmgpu_update_phase(ecmgpu_sim* sim, int phase);
/* Strip boundaries along x: n_ranks+1 ascending values; rank r owns [bounds[r], bounds[r+1])
 * (the first and last strip extend to infinity).  halo_width: agents this close to a border are
 * mirrored to the neighbour; interior strips must be at least this wide.  Agents outside this
 * rank's strip are deactivated here (another rank owns them).
 * Re-balancing: call it again with new borders after writing the GLOBAL state (positions, velocities,
 * preferred velocities, attraction points, forces, active flags, replan-pending flags of every agent) to
 * the slot arrays of every rank; message buffers, peer mappings and sequence numbers are kept, and the call
 * fails with ECMGPU_ERR_CAPACITY if the new borders need larger messages than the first call allocated. */
int ecmgpu_comm_set_strips(ecmgpu_sim* sim, const float* bounds, float halo_width);

#ifdef __cplusplus
}
#endif
#endif /* ECM_B200_H */
