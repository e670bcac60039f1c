/* ecm_b200_host.h - C ABI of the host-side (CPU, C++17) helpers that sit either side of the GPU tick:
 * ECM construction for lattice worlds and flattening into the structure-of-arrays layout the GPU
 * library (ecm_b200.h) consumes, plus global path planning.  None of this is on the per-tick hot path.
 *
 * Reference interfaces restated here (all under /root/reference):
 *   ECMGenerator::GenerateECM            ECMGenerator/ECMGenerator.h:19      (lattice worlds only)
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
/* Wrap caller-provided flat arrays (copied). */
ecmhost_world* ecmhost_world_from_arrays(const ecmhost_world_view* view);
void ecmhost_world_free(ecmhost_world* w);
int ecmhost_world_get_view(const ecmhost_world* w, ecmhost_world_view* out);

#ifdef __cplusplus
}
#endif
#endif
