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

// Appends one polygonal obstacle (CCW vertex order) and fills next/prev/convex with the
// reference's rule (/root/reference/ECMGenerator/ECMDataTypes.cpp:23-61).
void AppendObstacle(FlatObstacles& o, const float* xy, int n);

}  // namespace ecmb200
