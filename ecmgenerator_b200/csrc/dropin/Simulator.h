// Host-side drop-in for the reference's ECM::Simulation::Simulator
// (/root/reference/ECMAgentSimulator/Simulator.h:59-188): same public method names, argument
// meaning, slot allocation order and error behaviour; the per-tick systems run on the GPU through
// the C ABI of include/ecm_b200.h.  C++17, no CUDA headers needed to use it.
//
// What stays on the host, exactly like the reference: the free-slot stack (Simulator.h:66-69), spawn areas drawing from
// C rand() in the reference's order (Area.h:39-53, Simulator.cpp:494-536; their validity tests run on the GPU in batches,
// see SetSpawnMode), global path planning on spawn / replan (Simulator.cpp:97-124) through ecmb200::PathPlanner.
// ValidSpawnLocation as a public query keeps the reference's O(N) scan over the host mirrors (Simulator.cpp:295-311).
//
// Differences from the reference, all at points where the reference has undefined behaviour:
//   * a failed path query keeps the agent's previous path (the reference stores a 0-point path and
//     later reads path.x[-1], Simulator.cpp:112-118, :554); SpawnAgent with an unplannable start /
//     goal returns -1 and does not consume a slot;
//   * construction takes the flattened world (csrc/host/flat_world.h) instead of ECM* /
//     Environment* - INTEGRATION.md shows the 30-line flattening of the reference's own objects.
#pragma once
#include <map>
#include <stack>
#include <string>
#include <vector>

#include "../host/flat_world.h"
#include "../host/planner.h"
#include "sim_types.h"

struct ecmgpu_sim;

namespace ECM {
namespace Simulation {

class Simulator;

// Stand-in for the reference's KDTree (KDTree.h:57-80) behind Simulator::GetKDTree(): there is no tree on this
// side - the neighbour structure is the per-tick uniform grid in HBM - so Construct / Clear are no-ops and
// KNearestAgents answers with the exact 5-NN of the CURRENT state (Simulator::FindNNearestNeighbors).
class KDTree {
public:
    void Construct(Simulator*) {}
    void Clear() {}
    void KNearestAgents(Simulator* simulation, int agent, int k, std::vector<Entity>& outAgents, int& outNumNeighbors);
};

class Simulator {
public:
    // ---- what a frame loop reads: raw pointers into host mirrors indexed by slot, refreshed by Update()
    PositionComponent* GetPositionData() const { return xy_; }
    VelocityComponent* GetVelocityData() const { return vel_; }
    VelocityComponent* GetPreferredVelocityData() const { return pref_vel_; }
    PositionComponent* GetAttractionPointData() const { return attraction_; }
    ClearanceComponent* GetClearanceData() const { return radius_; }
    PathComponent* GetPathData() const { return routes_; }
    bool* GetActiveFlags() const { return alive_; }
    int GetNumAgents() const { return count_; }
    int GetLastIndex() const { return last_slot_; }
    float GetSimulationStepTime() const { return tick_seconds_; }
    ecmb200::PathPlanner* GetECMPathPlanner() { return planner_; }
    KDTree* GetKDTree() const { return const_cast<KDTree*>(&kdtree_); }
    const ecmb200::FlatWorld* GetEnvironment() const { return world_; }        // the flattened Environment + ECM
    const ecmb200::FlatObstacles& GetObstacles() const { return obstacles_; }  // world obstacles + AddObstacleArea boxes
    ecmgpu_sim* GetGpuHandle() const { return gpu_; }
    const std::string& LastError() const { return error_; }
    // debug hooks of the reference's renderer (Simulator.h:131-132): the neighbours of agent NN_TO_DRAW, kept by the query
    int NN_TO_DRAW = 0;
    std::vector<int> NEAREST_NEIGHBORS;

    // ---- life cycle.  `w` and `routePlanner` are borrowed, like ECM* / ECMPathPlanner* / Environment* in the reference
    Simulator(const ecmb200::FlatWorld* w, ecmb200::PathPlanner* routePlanner, int maxAgents, float simStepTime, int device = 0);
    ~Simulator();
    void Initialize();           // host mirrors + the GPU simulator; throws std::runtime_error without a CUDA device
    void Update(float ignored);  // the constructor's step is used, like the reference (Simulator.cpp:314-323)
    void Reset();

    // ---- how the spawn areas run (Simulator::UpdateSpawnAreas, Simulator.cpp:494-536)
    //   SPAWN_RAND_BATCHED (default): C rand() consumed in exactly the reference's order - same slots, same positions -
    //       but the ValidSpawnLocation tests of a tick go to the GPU in batches (ecmgpu_valid_spawn_locations on the
    //       neighbour grid) instead of one O(N) host scan per attempt.  The draws of a batch assume every first attempt
    //       succeeds; where one does not, the generator is rewound to that point (glibc: rand()'s state is switchable,
    //       setstate(3)) and that request is finished attempt by attempt.  Without glibc it degrades to SEQUENTIAL.
    //   SPAWN_RAND_SEQUENTIAL: the reference's loop as it is, one host scan per attempt.
    //   SPAWN_DEVICE_COUNTER: draws from a counter-based generator ON the device (ecmgpu_draw_spawns); reproducible for a
    //       seed, NOT the rand() stream - parity with the reference is statistical only.
    enum SpawnMode { SPAWN_RAND_BATCHED = 0, SPAWN_RAND_SEQUENTIAL = 1, SPAWN_DEVICE_COUNTER = 2 };
    void SetSpawnMode(SpawnMode mode, unsigned long long seed = 0) { spawn_mode_ = mode; spawn_seed_ = seed; }
    // spawn attempts answered by the GPU / by the host scan since construction (what the batching saved)
    long long SpawnChecksOnDevice() const { return spawn_checks_device_; }
    long long SpawnChecksOnHost() const { return spawn_checks_host_; }

    // ---- agents
    int SpawnAgent(const Point& from, const Point& to, float radius, float speed);
    void DestroyAgent(int slot);
    void UpdatePath(const Entity& slot, const Point& from, const Point& to);
    void AddPosition(Entity slot, float x, float y);
    bool ValidSpawnLocation(const Point& where, float radius) const;

    // ---- queries on the current state
    // exact 5-NN (DESIGN.md "Neighbour contract"); `out` must have size k == 5
    void FindNNearestNeighbors(const Entity& slot, int k, std::vector<Entity>& out, int& found);
    // the reference's unused brute-force variant (Simulator.cpp:228-257): same exact answer here, resizes `out`
    void FindNNearestNeighborsDeprecated(const Entity& slot, int k, std::vector<Entity>& out, int& found);
    // flat obstacle-vertex indices in (obstacle, vertex) order instead of ObstacleVertex*
    void FindNearestObstacles(const Entity& slot, float rangeSquared, std::vector<int>& out) const;

    // ---- areas (the editor's API, Command.cpp:33-159)
    int AddSpawnArea(const Point& centre, const Vec2& halfExtent, const SpawnConfiguration& profile, int ID = -1);
    int AddGoalArea(const Point& centre, const Vec2& halfExtent, int ID = -1);
    // Simulator.cpp:383-412: a box obstacle for FindNearestObstacles / ORCA.  Like the reference the area itself is
    // not recorded (GetObstacleAreas() stays empty, the returned ID is 0) and paths are not replanned; updateECM = true
    // would need the host-side ECM generator (out of scope here): the obstacle is added, LastError() says so.
    int AddObstacleArea(const Point& centre, const Vec2& halfExtent, bool updateECM = false);
    void RemoveArea(SimAreaType kind, int ID);
    void ConnectSpawnGoalAreas(int spawnID, int goalID, float agentsPerSecond = 0.0f);
    void DeconnectSpawnGoalAreas(int spawnID, int goalID);
    std::vector<int> GetConnectedAreas(int sourceID, SimAreaType kind);
    SpawnArea* GetSpawnArea(int ID);
    GoalArea* GetGoalArea(int ID);
    std::map<int, SpawnArea>& GetSpawnAreas() { return spawn_areas_; }
    std::map<int, GoalArea>& GetGoalAreas() { return goal_areas_; }
    std::vector<ObstacleArea>& GetObstacleAreas() { return obstacle_areas_; }

private:
    void ReleaseAll();
    void TrimLastSlot();
    void RunSpawnAreas();
    void RunSpawnAreasSequential();
    struct SpawnRequest { SpawnArea* area; int goalArea; float clearance, speed; };
    void RunSpawnRequestsBatched(const std::vector<SpawnRequest>& req);
    void RunSpawnRequestsOnDevice(const std::vector<SpawnRequest>& req);
    int SpawnChecked(const Point& from, const Point& to, float radius, float speed);  // SpawnAgent after a validity test already made
    bool ClashesWithThisTick(const Point& p, float clearance) const;
    void StoreRoute(int slot, const std::vector<ecmb200::P2f>& polyline);
    void Check(int rc, const char* what);

    const ecmb200::FlatWorld* world_;
    ecmb200::PathPlanner* planner_;
    ecmgpu_sim* gpu_ = nullptr;
    int device_;

    int capacity_;
    int count_ = 0;
    std::stack<int> free_slots_;
    bool* alive_ = nullptr;
    int last_slot_ = -1;
    float tick_seconds_;

    std::map<int, SpawnArea> spawn_areas_;
    std::map<int, GoalArea> goal_areas_;
    std::vector<ObstacleArea> obstacle_areas_;
    ecmb200::FlatObstacles obstacles_;  // what the GPU holds: the world's obstacles, then the boxes added at run time
    KDTree kdtree_;
    int next_spawn_id_ = 0;
    int next_goal_id_ = 0;

    PositionComponent* xy_ = nullptr;
    PositionComponent* attraction_ = nullptr;
    VelocityComponent* pref_vel_ = nullptr;
    VelocityComponent* vel_ = nullptr;
    ClearanceComponent* radius_ = nullptr;
    SpeedComponent* pref_speed_ = nullptr;
    PathComponent* routes_ = nullptr;

    SpawnMode spawn_mode_ = SPAWN_RAND_BATCHED;
    unsigned long long spawn_seed_ = 0, spawn_counter_ = 0;
    long long spawn_checks_device_ = 0, spawn_checks_host_ = 0;
    std::vector<Point> spawned_this_tick_;  // agents spawned since the device positions the batch tests ran on

    bool nbr_valid_ = false;
    std::vector<int> nbr_ids_, nbr_counts_;
    std::vector<int> event_scratch_;
    std::string error_;
};

}  // namespace Simulation
}  // namespace ECM
