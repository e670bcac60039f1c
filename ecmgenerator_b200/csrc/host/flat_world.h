// Flattened, structure-of-arrays description of an ECM navigation mesh plus its ORCA obstacles.
// This is the interchange format between host-side ECM construction and the GPU tick
// (SURVEY.md Appendix A).  Conventions follow the reference data model:
//   ECMVertex/ECMEdge/ECMHalfEdge   /root/reference/ECMGenerator/ECM.h:20-50
//   half-edge id = 2*edge + side    /root/reference/ECMGenerator/ECM.h:69
//   cell id      = 2*edge + side    /root/reference/ECMGenerator/ECMCellCollection.cpp:17-45
//   ObstacleVertex{p,prev,next,isConvex} /root/reference/ECMGenerator/ECMDataTypes.h:160-166
#pragma once
#include <vector>
#include <cstdint>

namespace ecmb200 {

struct FlatECM {
    // vertices
    std::vector<float> vert_xy;     // 2*nV
    std::vector<float> vert_clear;  // nV   distance to the closest obstacle point
    std::vector<int>   vert_he;     // nV   one outgoing half-edge
    // edges, stored once, directed v0 -> v1
    std::vector<int>   edge_v;      // 2*nE (v0, v1)
    // closest obstacle points: L0 R0 L1 R1 (x,y each) = left/right site point at v0 / at v1,
    // "left" relative to the direction v0->v1.  In reference terms:
    //   half_edges[0] = {target v1, closest_left = L0, closest_right = R0}
    //   half_edges[1] = {target v0, closest_left = R1, closest_right = L1}
    std::vector<float> edge_cl;     // 8*nE
    std::vector<int>   he_next;     // 2*nE  next outgoing half-edge around the same source vertex
    int num_vertices() const { return (int)vert_clear.size(); }
    int num_edges() const { return (int)edge_v.size() / 2; }
};

struct FlatObstacles {
    std::vector<float>   xy;        // 2*nO   obstacle vertices, obstacle-major, CCW per obstacle
    std::vector<int>     next;      // nO
    std::vector<int>     prev;      // nO
    std::vector<uint8_t> convex;    // nO
    std::vector<int>     first;     // nObst+1  offsets of each obstacle's vertex run
    int num_vertices() const { return (int)next.size(); }
    int num_obstacles() const { return first.empty() ? 0 : (int)first.size() - 1; }
};

struct FlatWorld {
    float bbox[4] = {0, 0, 0, 0};   // xmin, ymin, xmax, ymax of the walkable area
    FlatECM ecm;
    FlatObstacles obst;
};

// Levels at which the reference's even-odd test (strict y comparisons, UtilityFunctions.cpp:68-83) miscounts for the cell
// polygon with vertex heights y[0..3] (polygon order): a maximal run of consecutive vertices exactly at one y whose
// neighbours before and after it lie on OPPOSITE sides is a real crossing of the ray that neither incident edge counts,
// so a point level with it and to its left gets the wrong parity - it can be "inside" a cell far to its right.  Touching
// runs (same side) and extremal vertices are counted correctly.  An axis-aligned lattice has no such level (its cells
// meet every vertex level from one side); a turned world has about two per cell.  Shared by the host locator
// (planner.cpp) and the device bins (csrc/ecmgpu.cu).
inline void pass_through_levels(const float y[4], std::vector<float>& out) {
    for (int i = 0; i < 4; i++) {
        const float L = y[i];
        if (!(L == L)) continue;
        if (y[(i + 3) & 3] == L) continue;  // not the first vertex of its run
        int j = i, len = 1;
        while (len < 4 && y[(j + 1) & 3] == L) { j++; len++; }
        if (len == 4) continue;             // all four on one level: the strict test never counts anything
        const float before = y[(i + 3) & 3], after = y[(j + 1) & 3];
        if ((before < L) != (after < L)) out.push_back(L == 0.0f ? 0.0f : L);
    }
}

// Appends one polygonal obstacle (CCW vertex order) and fills next/prev/convex with the
// reference's rule (/root/reference/ECMGenerator/ECMDataTypes.cpp:23-61).
void AppendObstacle(FlatObstacles& o, const float* xy, int n);

}  // namespace ecmb200
