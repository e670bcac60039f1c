// Device views of the static world (flattened ECM + obstacles + the static bin grid) and of the
// per-tick neighbour grid.  All pointers are device pointers owned by ecmgpu_sim.
#pragma once
#include "geom.cuh"

namespace ecm {

// Flattened ECMGraph (reference: ECM.h:20-50).  edge_cl holds L0 R0 L1 R1 per edge:
//   half_edges[0].closest_left = L0, .closest_right = R0, half_edges[1].closest_left = R1, .closest_right = L1.
struct EcmView {
    int n_vertices, n_edges;
    const float2* vert_xy;   // [nV]
    const int2* edge_v;      // [nE] (v0, v1)
    const float2* edge_cl;   // [4*nE]
};

// Obstacle vertices in (obstacle, vertex) order (reference: ECMDataTypes.h:160-166).
struct ObstView {
    int n;
    const float2* xy;
    const int* next;
    const int* prev;
    const unsigned char* convex;
    // unit direction of segment i -> next[i], precomputed on the host with the same IEEE sqrt / divide as
    // Vec2::Normalize (bit-identical to normalising on the fly; ORCA.cpp:88, :111-112, :226-227, :234-235)
    const float2* dir;
};

// Uniform bins over the walkable-area bbox (+margin).  Per bin two ascending id lists (CSR):
//   cell list  - ECM cells whose bounding box touches the bin: scanning it in order reproduces
//                "first containing cell in index order" of PointLocationQueryLinear
//                (ECMCellCollection.cpp:57-90);
//   obst list  - obstacle segments within max_range of the bin: scanning it in order reproduces the
//                (obstacle, vertex) order of FindNearestObstacles (Simulator.cpp:259-292).
// A query point outside the grid falls back to the exhaustive scans (exactness over speed).
struct BinView {
    float x0, y0, inv_bin;
    int w, h;
    const int* cell_start;   // [w*h+1]
    const int* cell_items;
    const int* obst_start;   // [w*h+1]
    const int* obst_items;
    // The reference's even-odd test (UtilityFunctions.cpp:54-86) uses strict y comparisons, so a ray that passes exactly
    // through a polygon vertex where the boundary crosses from below to above (a "pass-through" vertex, not a y-extremum)
    // counts that crossing zero times: a point LEVEL with such a vertex and to its left gets the wrong parity and can be
    // reported inside a cell arbitrarily far to its right (golden scene oblique_small).  The bbox argument behind the
    // per-bin lists does not hold for those points.  They are recognised exactly - level_y is the sorted set of the
    // pass-through vertex heights (host/flat_world.h pass_through_levels; empty for axis-aligned lattices, whose paths
    // DO run along vertex levels), level_bits a hash bitmap in front of it - and answered from row_items: per bin ROW
    // the ascending list of the cells whose y-extent reaches into the row (a cell containing p under the reference's
    // predicate has an edge with min y < p.y < max y, so it is in the list of p's row).
    // 1: every ECM cell lies inside the grid and every obstacle vertex at least max_range inside it (checked on the host
    // when the lists are built): a point OUTSIDE the grid is then in no cell and has no obstacle in range, and the
    // exhaustive scans that would answer it exactly are skipped (an agent pushed out of the world would otherwise cost
    // one thread a scan of every cell and every segment, every tick)
    int closed;
    const float* level_y;
    const unsigned* level_bits;
    int n_levels, level_shift;
    const int* row_start;    // [h+1]
    const int* row_items;
    __device__ __forceinline__ bool level_hit(float y) const {
        if (n_levels == 0) return false;
        const unsigned u = __float_as_uint(y == 0.0f ? 0.0f : y);  // -0 == +0
        const unsigned hsh = (u * 2654435761u) >> level_shift;
        if (!((__ldg(&level_bits[hsh >> 5]) >> (hsh & 31u)) & 1u)) return false;
        int lo = 0, hi = n_levels - 1;
        while (lo <= hi) {
            const int mid = (lo + hi) >> 1;
            const float v = __ldg(&level_y[mid]);
            if (v == y) return true;
            if (v < y) lo = mid + 1;
            else hi = mid - 1;
        }
        return false;
    }
    __device__ __forceinline__ int row_of(v2 p) const {
        const float fy = (p.y - y0) * inv_bin;
        if (!(fy >= 0.0f) || !(fy < (float)h)) return -1;
        return (int)fy;
    }
    __device__ __forceinline__ int bin_of(v2 p) const {
        float fx = (p.x - x0) * inv_bin, fy = (p.y - y0) * inv_bin;
        // !(>=) also rejects NaN
        if (!(fx >= 0.0f) || !(fy >= 0.0f) || !(fx < (float)w) || !(fy < (float)h)) return -1;
        return (int)fy * w + (int)fx;
    }
};

// Per-tick neighbour grid: agents active at the start of the tick, counting-sorted by cell key
// (row-major), with a structure-of-arrays snapshot of their pre-tick state in sorted order.
struct GridView {
    float x0, y0, cell, inv_cell;
    int w, h;
    int n_sorted;             // number of agents in the snapshot
    const int* cell_start;    // [w*h+1] exclusive scan of per-cell counts
    const float2* s_pos;      // [n_sorted] pre-tick position
    const float2* s_vel;      // [n_sorted] pre-tick velocity
    const float* s_rad;       // [n_sorted] radius
    const int* s_slot;        // [n_sorted] internal agent index (what the per-agent arrays are indexed by)
    // The library renumbers agents in spatial order when they are loaded (ecmgpu.cu "spatial renumbering"): ext_of maps an
    // internal index to the caller's slot id, nullptr = identity.  The reference's tie-break is by SLOT id.
    const int* ext_of;
    __device__ __forceinline__ int slot_of_row(int row) const {
        const int i = __ldg(&s_slot[row]);
        return ext_of ? __ldg(&ext_of[i]) : i;
    }
    __device__ __forceinline__ void cell_of(v2 p, int& cx, int& cy) const {
        float fx = (p.x - x0) * inv_cell, fy = (p.y - y0) * inv_cell;
        // clamp (NaN -> 0): agents outside the grid live in the border cells
        cx = fx >= 0.0f ? (fx < (float)w ? (int)fx : w - 1) : 0;
        cy = fy >= 0.0f ? (fy < (float)h ? (int)fy : h - 1) : 0;
    }
};

}  // namespace ecm
