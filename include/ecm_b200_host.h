/* ecm_b200_host.h - C ABI of the host-side (CPU, C++17) helpers that sit either side of the GPU tick:
 * ECM construction for lattice worlds and flattening into the structure-of-arrays layout the GPU
 * library (ecm_b200.h) consumes, plus global path planning.  None of this is on the per-tick hot path.
 *
 * Reference interfaces restated here (all under /root/reference):
 *   ECMGenerator::GenerateECM            ECMGenerator/ECMGenerator.h:19      (lattice worlds in closed form; polygonal scenes directly)
 *   Environment::AddWalkableArea/AddObstacle  ECMGenerator/Environment.h:47-48
 *   ECMPathPlanner::FindPath             ECMGenerator/ECMPathPlanner.h:55
 */
#ifndef ECM_B200_HOST_H
#define ECM_B200_HOST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ecmhost_world ecmhost_world;

/* Flat views into a world; pointers stay valid until ecmhost_world_free(). Layout: see
 * ecmgenerator_b200/csrc/host/flat_world.h. */
typedef struct ecmhost_world_view {
    float bbox[4];             /* xmin ymin xmax ymax */
    int n_vertices, n_edges, n_obst_vertices, n_obstacles;
    const float* vert_xy;      /* 2*n_vertices */
    const float* vert_clear;   /* n_vertices */
    const int* vert_he;        /* n_vertices */
    const int* edge_v;         /* 2*n_edges */
    const float* edge_cl;      /* 8*n_edges: L0 R0 L1 R1 */
    const int* he_next;        /* 2*n_edges */
    const float* obst_xy;      /* 2*n_obst_vertices */
    const int* obst_next;      /* n_obst_vertices */
    const int* obst_prev;      /* n_obst_vertices */
    const uint8_t* obst_convex;/* n_obst_vertices */
    const int* obst_first;     /* n_obstacles+1 */
} ecmhost_world_view;

/* Blocks bx[0..nbx) x by[0..nby) separated by streets of width W; NULL on invalid input. */
ecmhost_world* ecmhost_lattice_world(int nbx, const float* bx, int nby, const float* by, float W,
                                     float x0, float y0);
/* ECMGenerator::GenerateECM (ECMGenerator/ECMGenerator.cpp:235-256) for a rectangular walkable area with polygonal
 * obstacles, without Boost: the segment Voronoi diagram of the area's edges and the obstacle edges is computed directly
 * (csrc/host/polygon_world.h) and its primary edges in free space become the ECM, in the reference's conventions.
 * poly_first[n_polys+1] indexes the counter-clockwise vertices in poly_xy; obstacles lie strictly inside the area and do
 * not touch.  For scenes of up to a few hundred edges (the reference's own test environments, Environment.cpp:27-185).
 * NULL on invalid input, with a message in `error` (may be NULL). */
ecmhost_world* ecmhost_polygon_world(const float bbox[4], int n_polys, const int* poly_first, const float* poly_xy, char* error, int error_cap);
/* Wrap caller-provided flat arrays (copied). */
ecmhost_world* ecmhost_world_from_arrays(const ecmhost_world_view* view);
void ecmhost_world_free(ecmhost_world* w);
int ecmhost_world_get_view(const ecmhost_world* w, ecmhost_world_view* out);

/* -- global path planning: ECMPathPlanner::FindPath (ECMGenerator/ECMPathPlanner.cpp:22-136) as
 * Simulator::UpdatePath calls it (ECMAgentSimulator/Simulator.cpp:97-124), batched over `threads`
 * host threads (0 = all).  A failed query yields a zero-length polyline. */
typedef struct ecmhost_paths ecmhost_paths;
ecmhost_paths* ecmhost_plan_paths(const ecmhost_world* w, int n, const float* start_xy, const float* goal_xy,
                                  const float* clearance, int threads);
int ecmhost_paths_count(const ecmhost_paths* p);        /* n */
int ecmhost_paths_succeeded(const ecmhost_paths* p);    /* queries that produced a path */
const int* ecmhost_paths_offsets(const ecmhost_paths* p); /* n+1 offsets in points */
const float* ecmhost_paths_xy(const ecmhost_paths* p);  /* 2 * offsets[n] floats */
void ecmhost_paths_free(ecmhost_paths* p);
/* ECMGraph::FindCell (ECMGenerator/ECM.cpp:191-194) on the host, for n points. */
int ecmhost_find_cells(const ecmhost_world* w, int n, const float* xy, int* out_cell);

#ifdef __cplusplus
}
#endif
#endif
