// Host-side global path planner over a flattened ECM.
//
// This is the caller of the hot path on the spawn / replan side: it produces the
// per-agent indicative route (polyline) the GPU's IRM kernel follows.  Behaviour restates the
// reference pipeline so that both sides plan identical polylines from identical queries:
//   ECMPathPlanner::FindPath          /root/reference/ECMGenerator/ECMPathPlanner.cpp:22-136
//   AStar::FindPath / ConstructPath   /root/reference/ECMGenerator/AStar.cpp:44-160, 184-212
//   CreateCorridor / ShrinkCorridor   /root/reference/ECMGenerator/ECMPathPlanner.cpp:146-214
//   TriangulateCorridor / SampleCorridorArc / FitPortalRange   ECMPathPlanner.cpp:216-320
//   Funnel                            /root/reference/ECMGenerator/ECMPathPlanner.cpp:323-411
// Differences that do not change results: point location goes through a bin index instead of the
// linear scan (same "lowest-index containing cell" answer), and the A* scratch arrays are reset
// through a touched-list instead of an O(V) sweep per query (AStar.cpp:162-176).
#pragma once
#include <vector>

#include "flat_world.h"

namespace ecmb200 {

struct P2f {
    float x, y;
};

// Lowest-index containing ECM cell, as ECMCellCollection::PointLocationQueryLinear
// (/root/reference/ECMGenerator/ECMCellCollection.cpp:57-90) returns it; -1 if none.
class CellLocator {
public:
    void Build(const FlatWorld& w, float bin = 0.0f);
    int FindCell(const FlatWorld& w, float x, float y) const;

private:
    float x0_ = 0, y0_ = 0, bin_ = 1;
    int w_ = 0, h_ = 0;
    std::vector<int> start_, items_;
    // A point exactly level with a cell vertex can be "inside" a cell far to its right under the reference's even-odd
    // test (strict y comparisons; csrc/device/world.cuh BinView::level_hit): such points are recognised through the
    // sorted set of vertex y values and answered from per-row cell lists instead of the bin list.
    std::vector<float> levels_;
    std::vector<int> row_start_, row_items_;
};

// ECM::RetractPoint (/root/reference/ECMGenerator/ECM.cpp:20-96) on a located cell.
bool RetractPoint(const FlatWorld& w, const CellLocator& loc, P2f p, P2f& out, int& out_edge);

class PathPlanner {
public:
    explicit PathPlanner(const FlatWorld* world);
    // Scratch state is per instance: use one PathPlanner per thread (they share `world` and `locator`).
    PathPlanner(const FlatWorld* world, const CellLocator* shared_locator);
    ~PathPlanner();

    // ECMPathPlanner::FindPath with preferredAdditionalClearance = 0 (Simulator.cpp:108-112).
    // Returns false (and an empty path) where the reference returns false.
    bool FindPath(P2f start, P2f goal, float clearance, std::vector<P2f>& out_path);

    const CellLocator& locator() const { return *loc_; }

private:
    bool AStar(P2f startLoc, P2f goalLoc, int startEdge, int goalEdge, float clearance, std::vector<int>& outPath);

    const FlatWorld* w_;
    const CellLocator* loc_;
    CellLocator* owned_loc_ = nullptr;
    // A* scratch (AStarNode, AStar.h:18-28)
    std::vector<float> g_, f_;
    std::vector<int> parent_;
    std::vector<unsigned char> visited_;
    std::vector<int> touched_;
};

// Plans n paths with `threads` worker threads (0 = hardware concurrency).  out_off[n+1] / out_xy
// receive the polylines back to back; a failed query contributes zero points.  Returns #successes.
int PlanPaths(const FlatWorld& w, int n, const float* start_xy, const float* goal_xy, const float* clearance, int threads,
              std::vector<int>& out_off, std::vector<float>& out_xy);

}  // namespace ecmb200
