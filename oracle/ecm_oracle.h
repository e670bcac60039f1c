/* TEST INFRASTRUCTURE - NOT PRODUCT CODE.
 *
 * ecm_oracle: plain-C, single-threaded CPU restatement of the reference's per-tick agent update
 * (Simulator::Update, /root/reference/ECMAgentSimulator/Simulator.cpp:314-323) over flat arrays.
 * Every function in ecm_oracle.c cites the reference file:line it follows.  Only tests/,
 * __graft_entry__.smoke() and bench.py's CPU-baseline legs may build, load or call this; the
 * product (ecmgenerator_b200/, include/) never does.
 *
 * PINNING: the reference ships no tests or golden vectors, so this restatement is
 * pinned against outputs of the reference itself compiled here (oracle/_ref, see oracle/Makefile):
 * tests/test_oracle_vs_reference.py requires bit-identical state after every tick in both
 * neighbour modes, and tests/golden/ holds vectors generated from oracle/_ref by
 * tests/golden/make_golden.py for the GPU box, where /root/reference does not exist.
 *
 * Build: gcc -std=c11 -O2 -ffp-contract=off (no -march: FMA contraction changes results, P4).
 */
#ifndef ECM_ORACLE_H
#define ECM_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct eo_sim eo_sim;

enum { EO_KNN_REF_KDTREE = 0, EO_KNN_EXACT = 1 };

/* World arrays are copied.  Layout: ecmgenerator_b200/csrc/host/flat_world.h. */
eo_sim* eo_create(int nV, const float* vert_xy, const float* vert_clear, int nE, const int* edge_v,
                  const float* edge_cl, int nO, const float* obst_xy, const int* obst_next, const int* obst_prev,
                  const uint8_t* obst_convex, int max_agents, float step, int knn_mode);
void eo_destroy(eo_sim* s);

/* SpawnAgent's writes (Simulator.cpp:175-199) for n agents, lowest free slot first, without the
 * ValidSpawnLocation scan; paths are given (path_off[n+1] offsets into path_xy points).  Agents
 * with fewer than 2 path points are skipped (out_slots[i] = -1).  Returns #loaded. */
int eo_bulk_load(eo_sim* s, int n, const float* pos_xy, const float* radius, const float* speed, const int* path_off,
                 const float* path_xy, int* out_slots);
void eo_set_path(eo_sim* s, int slot, const float* xy, int n);
void eo_set_kinematics(eo_sim* s, int slot, float x, float y, float vx, float vy);
void eo_set_attraction(eo_sim* s, int slot, float x, float y);
void eo_destroy_agent(eo_sim* s, int slot);

/* One Simulator::Update (UpdateSpawnAreas excluded: no spawn areas).  Replans requested this tick
 * (Simulator.cpp:581-587) are NOT executed - the planner is host-side - but recorded, in slot order. */
void eo_step(eo_sim* s);
int eo_num_replans(const eo_sim* s);
const int* eo_replans(const eo_sim* s);
int eo_num_destroyed(const eo_sim* s);
const int* eo_destroyed(const eo_sim* s);

int eo_num_agents(const eo_sim* s);
int eo_last_index(const eo_sim* s);
void eo_get_state(const eo_sim* s, int count, float* pos, float* vel, float* prefvel, float* attraction, float* force,
                  uint8_t* active);

/* Piecewise queries on the current state (for unit parity). */
void eo_query_cells(const eo_sim* s, int n, const float* xy, int* out_cell);
void eo_retract(const eo_sim* s, int n, const float* xy, uint8_t* ok, float* out_xy, int* out_edge);
/* All active slots < count in ascending order with one shared, zero-initialised 5-entry cache (the
 * role of ORCA::m_NeighborCache, ORCA.h:100).  out_counts[i] = -1 for inactive slots. */
void eo_query_neighbors(eo_sim* s, int count, int* out_ids, int* out_counts);
int eo_query_obstacles(const eo_sim* s, int slot, int* out_ids, int cap);
/* ORCA::GetVelocity for one slot given its neighbour list. */
void eo_orca_velocity(const eo_sim* s, int slot, int n_neighbors, const int* neighbors, float* out_v);
/* counters accumulated over eo_step calls: [0] LP3D invocations, [1] LP calls, [2] max obstacle
 * neighbours seen, [3] IRM failures (replans), [4] point-location failures, [5] agent-updates,
 * [6] obstacle segments with a concave end vertex handed to GenerateConstraints, [7] oblique ones */
const long long* eo_counters(const eo_sim* s);

/* test hook: the libstdc++ std::sort restatement used by the KD-tree build, on its own */
void eo_test_std_sort(int* first, int n, const float* pos_xy, int axis);

#ifdef __cplusplus
}
#endif
#endif
