// Host-side drop-in for the reference's ECM::Simulation::Simulator
// (/root/reference/ECMAgentSimulator/Simulator.h:59-188): same public method names, argument
// meaning, slot allocation order and error behaviour; the per-tick systems run on the GPU through
// the C ABI of include/ecm_b200.h.  C++17, no CUDA headers needed to use it.
//
// What stays on the host, exactly like the reference: the free-slot stack (Simulator.h:66-69),
// ValidSpawnLocation's O(N) scan (Simulator.cpp:295-311), spawn areas with C rand()
// (Area.h:39-53, Simulator.cpp:494-536), global path planning on spawn / replan
// (Simulator.cpp:97-124) through ecmb200::PathPlanner.
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

struct ecmgpu_sim;

namespace ECM {

struct Point {
    float x = 0, y = 0;
    Point() {}
    Point(float x_, float y_) : x(x_), y(y_) {}
};
struct Vec2 {
    float x = 0, y = 0;
    Vec2() {}
    Vec2(float x_, float y_) : x(x_), y(y_) {}
};

namespace Simulation {

typedef int Entity;

// component structs: identical layout to Simulator.h:29-57
struct PositionComponent { float x; float y; };
struct VelocityComponent { float dx; float dy; };
struct ClearanceComponent { float clearance; };
struct SpeedComponent { float speed; };
struct PathComponent { int currentIndex; int numPoints; float* x; float* y; };

enum SimAreaType { NONE, WALKABLE, SPAWN, GOAL, OBSTACLE };  // Area.h:11-18

struct Area {  // Area.h:31-75
    int ID = 0;
    Point Position;
    float HalfHeight = 0;
    float HalfWidth = 0;
    SimAreaType Type = NONE;
    Point GetRandomPositionInArea();
    bool Intersects(const Point position) const;
};
struct GoalArea : public Area { GoalArea() { Type = GOAL; } };
struct SpawnConfiguration {  // Area.h:82-88
    float preferredSpeedMin = 5.0f, preferredSpeedMax = 10.0f, clearanceMin = 5.0f, clearanceMax = 10.0f;
};
struct SpawnArea : public Area {
    SpawnArea() { Type = SPAWN; }
    SpawnConfiguration spawnConfiguration;
    std::vector<int> connectedGoalAreas;
    std::vector<float> spawnRate;
    std::vector<float> timeSinceLastSpawn;
};

struct ObstacleArea : public Area {  // Area.h:101-105
    ObstacleArea() { Type = OBSTACLE; }
    std::vector<Point> obstacleVerts;
};

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
    // `world` and `planner` are borrowed, like ECM* / ECMPathPlanner* / Environment* in the reference.
    Simulator(const ecmb200::FlatWorld* world, ecmb200::PathPlanner* planner, int maxAgents, float simStepTime, int device = 0);
    ~Simulator();

    int SpawnAgent(const Point& start, const Point& goal, float clearance, float preferredSpeed);
    void DestroyAgent(int idx);

    void Initialize();      // allocates host mirrors and the GPU simulator; throws std::runtime_error without a CUDA device
    void Update(float dt);  // dt is ignored, the constructor's step is used (Simulator.cpp:314-323)
    void Reset();

    void AddPosition(Entity entity, float x, float y);

    int AddSpawnArea(const Point& position, const Vec2& halfSize, const SpawnConfiguration& config, int ID = -1);
    int AddGoalArea(const Point& position, const Vec2& halfSize, int ID = -1);
    // Simulator.cpp:383-412: a box obstacle for FindNearestObstacles / ORCA.  Like the reference the area itself is
    // not recorded (GetObstacleAreas() stays empty, the returned ID is 0) and paths are not replanned; updateECM = true
    // would need the host-side ECM generator (out of scope here): the obstacle is added, LastError() says so.
    int AddObstacleArea(const Point& position, const Vec2& halfSize, bool updateECM = false);
    std::vector<ObstacleArea>& GetObstacleAreas() { return m_ObstacleAreas; }
    void RemoveArea(SimAreaType areaType, int ID);
    void ConnectSpawnGoalAreas(int spawnID, int goalID, float spawnRate = 0.0f);
    void DeconnectSpawnGoalAreas(int spawnID, int goalID);
    std::map<int, SpawnArea>& GetSpawnAreas() { return m_SpawnAreas; }
    std::map<int, GoalArea>& GetGoalAreas() { return m_GoalAreas; }
    SpawnArea* GetSpawnArea(int ID);
    GoalArea* GetGoalArea(int ID);
    std::vector<int> GetConnectedAreas(int sourceID, SimAreaType type);

    // exact 5-NN (DESIGN.md "Neighbour contract"); outNeighbors must have size n == 5
    void FindNNearestNeighbors(const Entity& agent, int n, std::vector<Entity>& outNeighbors, int& outNNeighbors);
    // the reference's unused brute-force variant (Simulator.cpp:228-257): same exact answer here, resizes outNeighbors
    void FindNNearestNeighborsDeprecated(const Entity& agent, int n, std::vector<Entity>& outNeighbors, int& outNNeighbors);
    // flat obstacle-vertex indices in (obstacle, vertex) order instead of ObstacleVertex*
    void FindNearestObstacles(const Entity& agent, float rangeSquared, std::vector<int>& outObstacles) const;
    bool ValidSpawnLocation(const Point& location, float clearance) const;
    void UpdatePath(const Entity& e, const Point& location, const Point& goal);

    // getters: raw pointers into host mirrors indexed by slot, refreshed by Update()
    int GetNumAgents() const { return m_NumEntities; }
    int GetLastIndex() const { return m_LastEntityIdx; }
    PositionComponent* GetPositionData() const { return m_Positions; }
    VelocityComponent* GetVelocityData() const { return m_Velocities; }
    VelocityComponent* GetPreferredVelocityData() const { return m_PreferredVelocities; }
    PathComponent* GetPathData() const { return m_Paths; }
    ClearanceComponent* GetClearanceData() const { return m_Clearances; }
    PositionComponent* GetAttractionPointData() const { return m_AttractionPoints; }
    bool* GetActiveFlags() const { return m_ActiveAgents; }
    ecmb200::PathPlanner* GetECMPathPlanner() { return m_Planner; }
    KDTree* GetKDTree() const { return const_cast<KDTree*>(&m_KDTree); }
    const ecmb200::FlatWorld* GetEnvironment() const { return m_World; }  // the flattened Environment + ECM
    const ecmb200::FlatObstacles& GetObstacles() const { return m_Obst; } // world obstacles + AddObstacleArea boxes
    float GetSimulationStepTime() const { return m_SimStepTime; }
    const std::string& LastError() const { return m_Error; }
    ecmgpu_sim* GetGpuHandle() const { return m_Gpu; }

    int NN_TO_DRAW = 0;
    std::vector<int> NEAREST_NEIGHBORS;

private:
    void ClearSimulator();
    void UpdateMaxAgentIndex();
    void UpdateSpawnAreas();
    void SetPathComponent(int e, const std::vector<ecmb200::P2f>& path);
    void Check(int rc, const char* what);

    const ecmb200::FlatWorld* m_World;
    ecmb200::PathPlanner* m_Planner;
    ecmgpu_sim* m_Gpu = nullptr;
    int m_Device;

    int m_MaxNumEntities;
    int m_NumEntities = 0;
    std::stack<int> m_freeEntitySpaces;
    bool* m_ActiveAgents = nullptr;
    int m_LastEntityIdx = -1;
    float m_SimStepTime;

    std::map<int, SpawnArea> m_SpawnAreas;
    std::map<int, GoalArea> m_GoalAreas;
    std::vector<ObstacleArea> m_ObstacleAreas;
    ecmb200::FlatObstacles m_Obst;  // what the GPU holds: the world's obstacles, then the boxes added at run time
    KDTree m_KDTree;
    int m_NextSpawnID = 0;
    int m_NextGoalID = 0;

    PositionComponent* m_Positions = nullptr;
    PositionComponent* m_AttractionPoints = nullptr;
    VelocityComponent* m_PreferredVelocities = nullptr;
    VelocityComponent* m_Velocities = nullptr;
    ClearanceComponent* m_Clearances = nullptr;
    SpeedComponent* m_PreferredSpeed = nullptr;
    PathComponent* m_Paths = nullptr;

    bool m_NeighborsValid = false;
    std::vector<int> m_NeighborIds, m_NeighborCounts;
    std::vector<int> m_EventScratch;
    std::string m_Error;
};

}  // namespace Simulation
}  // namespace ECM
