#include "Simulator.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

#include "../../../include/ecm_b200.h"

// Float semantics: compiled with -ffp-contract=off like the rest of the host library.

namespace ECM {
namespace Simulation {

// Area::GetRandomPositionInArea (Area.h:39-53): three rand() draws, the first one unused.
Point Area::GetRandomPositionInArea() {
    float r = static_cast<float>(rand()) / static_cast<float>(RAND_MAX);
    (void)r;
    float xMin = Position.x - HalfWidth, xMax = Position.x + HalfWidth;
    float yMin = Position.y - HalfHeight, yMax = Position.y + HalfHeight;
    float randX = xMin + static_cast<float>(rand()) / (static_cast<float>(RAND_MAX / (xMax - xMin)));
    float randY = yMin + static_cast<float>(rand()) / (static_cast<float>(RAND_MAX / (yMax - yMin)));
    return Point(randX, randY);
}

bool Area::Intersects(const Point position) const {  // Area.h:66-74
    bool result = position.x <= (Position.x + HalfWidth);
    result &= position.x >= (Position.x - HalfWidth);
    result &= position.y <= (Position.y + HalfHeight);
    result &= position.y >= (Position.y - HalfHeight);
    return result;
}

Simulator::Simulator(const ecmb200::FlatWorld* world, ecmb200::PathPlanner* planner, int maxAgents, float simStepTime, int device)
    : m_World(world), m_Planner(planner), m_Device(device), m_MaxNumEntities(maxAgents), m_SimStepTime(simStepTime) {
    for (int i = maxAgents - 1; i >= 0; i--) m_freeEntitySpaces.push(i);  // Simulator.h:66-69
}

Simulator::~Simulator() { ClearSimulator(); }

void Simulator::Check(int rc, const char* what) {
    if (rc == ECMGPU_OK) return;
    m_Error = std::string(what) + ": " + (m_Gpu ? ecmgpu_last_error(m_Gpu) : ecmgpu_last_error(nullptr));
    throw std::runtime_error(m_Error);  // no CPU fallback: a GPU failure is fatal for the simulation
}

void Simulator::Initialize() {  // Simulator.cpp:21-60
    const int n = m_MaxNumEntities;
    m_LastEntityIdx = -1;
    m_ActiveAgents = new bool[n]();
    m_Positions = new PositionComponent[n]();
    m_AttractionPoints = new PositionComponent[n]();
    m_Velocities = new VelocityComponent[n]();
    m_PreferredVelocities = new VelocityComponent[n]();
    m_PreferredSpeed = new SpeedComponent[n]();
    m_Clearances = new ClearanceComponent[n]();
    m_Paths = new PathComponent[n];
    for (int i = 0; i < n; i++) {
        m_Paths[i].x = nullptr;
        m_Paths[i].y = nullptr;
        m_Paths[i].currentIndex = -1;
        m_Paths[i].numPoints = 0;
    }
    ecmgpu_params prm;
    memset(&prm, 0, sizeof(prm));
    prm.device = m_Device;
    prm.max_agents = n;
    prm.step = m_SimStepTime;
    prm.record_neighbors = 0;
    Check(ecmgpu_create(&prm, &m_Gpu), "ecmgpu_create");
    const auto& e = m_World->ecm;
    Check(ecmgpu_set_ecm(m_Gpu, m_World->bbox, e.num_vertices(), e.vert_xy.data(), e.vert_clear.data(), e.num_edges(),
                         e.edge_v.data(), e.edge_cl.data()), "ecmgpu_set_ecm");
    m_Obst = m_World->obst;
    const auto& o = m_Obst;
    Check(ecmgpu_set_obstacles(m_Gpu, o.num_vertices(), o.xy.data(), o.next.data(), o.prev.data(), o.convex.data()), "ecmgpu_set_obstacles");
    printf("SIMULATOR: Data for %d agents was created.\n", n);
}

void Simulator::ClearSimulator() {  // Simulator.cpp:62-95
    if (m_Paths) {
        for (int i = 0; i < m_MaxNumEntities; i++) {
            delete[] m_Paths[i].x;
            delete[] m_Paths[i].y;
        }
    }
    delete[] m_Positions; delete[] m_AttractionPoints; delete[] m_Velocities; delete[] m_PreferredVelocities;
    delete[] m_PreferredSpeed; delete[] m_Paths; delete[] m_ActiveAgents; delete[] m_Clearances;
    m_Positions = m_AttractionPoints = nullptr;
    m_Velocities = m_PreferredVelocities = nullptr;
    m_PreferredSpeed = nullptr; m_Paths = nullptr; m_ActiveAgents = nullptr; m_Clearances = nullptr;
    if (m_Gpu) {
        ecmgpu_destroy(m_Gpu);
        m_Gpu = nullptr;
        printf("SIMULATOR: Data was destroyed.\n");
    }
}

void Simulator::SetPathComponent(int e, const std::vector<ecmb200::P2f>& path) {  // Simulator.cpp:114-123
    PathComponent& pc = m_Paths[e];
    delete[] pc.x;
    delete[] pc.y;
    const int n = (int)path.size();
    pc.x = new float[n];
    pc.y = new float[n];
    pc.numPoints = n;
    pc.currentIndex = 0;
    std::vector<float> xy(2 * (size_t)n);
    for (int j = 0; j < n; j++) {
        pc.x[j] = xy[2 * j] = path[j].x;
        pc.y[j] = xy[2 * j + 1] = path[j].y;
    }
    Check(ecmgpu_set_path(m_Gpu, e, xy.data(), n), "ecmgpu_set_path");
}

// Simulator::UpdatePath (Simulator.cpp:97-124).  A failed query keeps the previous path.
void Simulator::UpdatePath(const Entity& e, const Point& location, const Point& goal) {
    std::vector<ecmb200::P2f> path;
    const bool ok = m_Planner->FindPath(ecmb200::P2f{location.x, location.y}, ecmb200::P2f{goal.x, goal.y}, m_Clearances[e].clearance, path);
    if (!ok || path.size() < 2) {
        if (m_Paths[e].numPoints >= 2) {  // re-arm the replan request instead of storing an unusable path
            std::vector<ecmb200::P2f> old((size_t)m_Paths[e].numPoints);
            for (int j = 0; j < m_Paths[e].numPoints; j++) old[j] = ecmb200::P2f{m_Paths[e].x[j], m_Paths[e].y[j]};
            SetPathComponent(e, old);
        }
        return;
    }
    SetPathComponent(e, path);
}

int Simulator::SpawnAgent(const Point& start, const Point& goal, float clearance, float preferredSpeed) {  // Simulator.cpp:168-200
    if (m_freeEntitySpaces.empty()) return -1;
    if (!ValidSpawnLocation(start, clearance)) return -1;
    std::vector<ecmb200::P2f> path;
    if (!m_Planner->FindPath(ecmb200::P2f{start.x, start.y}, ecmb200::P2f{goal.x, goal.y}, clearance, path) || path.size() < 2) return -1;
    m_NumEntities++;
    int idx = m_freeEntitySpaces.top();
    m_freeEntitySpaces.pop();
    m_LastEntityIdx = m_LastEntityIdx < idx ? idx : m_LastEntityIdx;
    m_Positions[idx].x = start.x;
    m_Positions[idx].y = start.y;
    m_Clearances[idx].clearance = clearance;
    m_PreferredSpeed[idx].speed = preferredSpeed;
    m_ActiveAgents[idx] = true;
    std::vector<float> xy(2 * path.size());
    for (size_t j = 0; j < path.size(); j++) { xy[2 * j] = path[j].x; xy[2 * j + 1] = path[j].y; }
    Check(ecmgpu_spawn(m_Gpu, idx, start.x, start.y, clearance, preferredSpeed, xy.data(), (int)path.size()), "ecmgpu_spawn");
    PathComponent& pc = m_Paths[idx];
    delete[] pc.x;
    delete[] pc.y;
    pc.numPoints = (int)path.size();
    pc.currentIndex = 0;
    pc.x = new float[path.size()];
    pc.y = new float[path.size()];
    for (size_t j = 0; j < path.size(); j++) { pc.x[j] = path[j].x; pc.y[j] = path[j].y; }
    m_PreferredVelocities[idx].dx = m_PreferredVelocities[idx].dy = 0.0f;
    m_Velocities[idx].dx = m_Velocities[idx].dy = 0.0f;
    m_AttractionPoints[idx].x = m_AttractionPoints[idx].y = 0.0f;
    m_NeighborsValid = false;
    return idx;
}

void Simulator::DestroyAgent(int idx) {  // Simulator.cpp:202-208
    m_NumEntities--;
    m_ActiveAgents[idx] = false;
    m_freeEntitySpaces.push(idx);
    Check(ecmgpu_destroy_agent(m_Gpu, idx), "ecmgpu_destroy_agent");
    m_NeighborsValid = false;
}

void Simulator::AddPosition(Entity entity, float x, float y) {  // Simulator.h:87-90
    m_Positions[entity].x = x;
    m_Positions[entity].y = y;
    Check(ecmgpu_write(m_Gpu, ECMGPU_POS, &m_Positions[entity], entity, 1), "ecmgpu_write");
    m_NeighborsValid = false;
}

bool Simulator::ValidSpawnLocation(const Point& location, float clearance) const {  // Simulator.cpp:295-311
    float clearanceSquared = clearance * clearance;
    for (int i = 0; i <= m_LastEntityIdx; i++) {
        if (!m_ActiveAgents[i]) continue;
        float dx = location.x - m_Positions[i].x, dy = location.y - m_Positions[i].y;
        float d = dx * dx + dy * dy;
        if (d < clearanceSquared) return false;
    }
    return true;
}

void Simulator::UpdateMaxAgentIndex() {  // Simulator.cpp:481-492
    int emptyCounter = 0;
    for (int i = m_LastEntityIdx; i >= 0; i--) {
        if (m_ActiveAgents[i]) break;
        emptyCounter++;
    }
    m_LastEntityIdx = m_LastEntityIdx - emptyCounter;
}

void Simulator::UpdateSpawnAreas() {  // Simulator.cpp:494-536
    for (auto iter = m_SpawnAreas.begin(); iter != m_SpawnAreas.end(); iter++) {
        SpawnArea& area = iter->second;
        for (int ga = 0; ga < (int)area.connectedGoalAreas.size(); ga++) {
            area.timeSinceLastSpawn[ga] += m_SimStepTime;
            int agentsToSpawn = area.timeSinceLastSpawn[ga] * area.spawnRate[ga];
            const int maxSpawnAttempts = 10;
            for (int i = 0; i < agentsToSpawn; i++) {
                float clearance = area.spawnConfiguration.clearanceMin;
                float speed = area.spawnConfiguration.preferredSpeedMin;
                bool foundValidLocation = false;
                Point start;
                for (int spawnAttempts = 0; spawnAttempts < maxSpawnAttempts; spawnAttempts++) {
                    start = area.GetRandomPositionInArea();
                    foundValidLocation = ValidSpawnLocation(start, clearance);
                    if (foundValidLocation) break;
                }
                if (!foundValidLocation) {
                    printf("Could not find a valid spawn position in the spawn area!\n");
                    continue;
                }
                Point goal = m_GoalAreas[area.connectedGoalAreas[ga]].GetRandomPositionInArea();
                SpawnAgent(start, goal, clearance, speed);
            }
            area.timeSinceLastSpawn[ga] -= (float)agentsToSpawn / area.spawnRate[ga];
        }
    }
}

void Simulator::Update(float /*dt*/) {  // Simulator.cpp:314-323
    UpdateMaxAgentIndex();
    UpdateSpawnAreas();
    m_NeighborsValid = false;
    if (m_LastEntityIdx < 0) return;
    Check(ecmgpu_update(m_Gpu), "ecmgpu_update");
    // events of this tick: arrivals (Simulator.cpp:564-566) then replans (Simulator.cpp:581-587), in slot
    // order like the reference's loop; the mirrors still hold the PRE-tick positions the replans start from
    const int count = m_LastEntityIdx + 1;
    m_EventScratch.resize(2 * (size_t)count);
    int nr = 0, nd = 0;
    Check(ecmgpu_poll_events(m_Gpu, m_EventScratch.data(), count, &nr, m_EventScratch.data() + count, count, &nd), "ecmgpu_poll_events");
    for (int k = 0; k < nd; k++) {
        const int e = m_EventScratch[count + k];
        m_NumEntities--;
        m_ActiveAgents[e] = false;
        m_freeEntitySpaces.push(e);
        // the reference assigns the goal as attraction point before destroying (Simulator.cpp:559-562)
        m_AttractionPoints[e].x = m_Paths[e].x[m_Paths[e].numPoints - 1];
        m_AttractionPoints[e].y = m_Paths[e].y[m_Paths[e].numPoints - 1];
    }
    for (int k = 0; k < nr; k++) {
        const int e = m_EventScratch[k];
        printf("Recalculate path...\n");
        Point cur(m_Positions[e].x, m_Positions[e].y);
        Point goal(m_Paths[e].x[m_Paths[e].numPoints - 1], m_Paths[e].y[m_Paths[e].numPoints - 1]);
        UpdatePath(e, cur, goal);
    }
    // refresh the mirrors the getters expose; destroyed agents keep their stale components (Appendix B.5)
    std::vector<PositionComponent> pos(count), att(count);
    std::vector<VelocityComponent> vel(count), pref(count);
    Check(ecmgpu_read(m_Gpu, ECMGPU_POS, pos.data(), 0, count), "ecmgpu_read");
    Check(ecmgpu_read(m_Gpu, ECMGPU_VEL, vel.data(), 0, count), "ecmgpu_read");
    Check(ecmgpu_read(m_Gpu, ECMGPU_PREFVEL, pref.data(), 0, count), "ecmgpu_read");
    Check(ecmgpu_read(m_Gpu, ECMGPU_ATTRACTION, att.data(), 0, count), "ecmgpu_read");
    for (int i = 0; i < count; i++) {
        if (!m_ActiveAgents[i]) continue;
        m_Positions[i] = pos[i];
        m_Velocities[i] = vel[i];
        m_PreferredVelocities[i] = pref[i];
        m_AttractionPoints[i] = att[i];
    }
}

void Simulator::Reset() {  // Simulator.cpp:325-333
    for (int i = 0; i <= m_LastEntityIdx; i++) {
        if (!m_ActiveAgents[i]) continue;
        DestroyAgent(i);
    }
}

int Simulator::AddSpawnArea(const Point& position, const Vec2& halfSize, const SpawnConfiguration& config, int ID) {  // Simulator.cpp:336-360
    SpawnArea sa;
    sa.HalfWidth = halfSize.x;
    sa.HalfHeight = halfSize.y;
    if (ID == -1) { sa.ID = m_NextSpawnID; m_NextSpawnID++; }
    else sa.ID = ID;
    sa.Position = position;
    sa.spawnConfiguration = config;
    m_SpawnAreas.emplace(sa.ID, sa);
    return sa.ID;
}

int Simulator::AddGoalArea(const Point& position, const Vec2& halfSize, int ID) {  // Simulator.cpp:362-381
    GoalArea ga;
    if (ID == -1) { ga.ID = m_NextGoalID; m_NextGoalID++; }
    else ga.ID = ID;
    ga.Position = position;
    ga.HalfHeight = halfSize.y;
    ga.HalfWidth = halfSize.x;
    m_GoalAreas.emplace(ga.ID, ga);
    return ga.ID;
}

void Simulator::RemoveArea(SimAreaType areaType, int ID) {  // Simulator.cpp:416-426
    if (areaType == SPAWN) m_SpawnAreas.erase(ID);
    if (areaType == GOAL) m_GoalAreas.erase(ID);
}

void Simulator::ConnectSpawnGoalAreas(int spawnID, int goalID, float spawnRate) {  // Simulator.cpp:428-440
    for (int ga : m_SpawnAreas[spawnID].connectedGoalAreas)
        if (ga == goalID) return;
    SpawnArea& sa = m_SpawnAreas[spawnID];
    sa.connectedGoalAreas.push_back(goalID);
    sa.timeSinceLastSpawn.push_back(0.0f);
    sa.spawnRate.push_back(spawnRate);
}

void Simulator::DeconnectSpawnGoalAreas(int spawnID, int goalID) {  // Simulator.cpp:442-460
    if (spawnID < 0 || spawnID >= (int)m_SpawnAreas.size()) return;
    auto& v = m_SpawnAreas[spawnID].connectedGoalAreas;
    for (size_t i = 0; i < v.size(); i++)
        if (v[i] == goalID) { v.erase(v.begin() + i); return; }
}

SpawnArea* Simulator::GetSpawnArea(int ID) {
    auto it = m_SpawnAreas.find(ID);
    return it != m_SpawnAreas.end() ? &it->second : nullptr;
}
GoalArea* Simulator::GetGoalArea(int ID) {
    auto it = m_GoalAreas.find(ID);
    return it != m_GoalAreas.end() ? &it->second : nullptr;
}

std::vector<int> Simulator::GetConnectedAreas(int sourceID, SimAreaType type) {  // Simulator.cpp:126-166
    if (type == SPAWN) {
        SpawnArea* sa = GetSpawnArea(sourceID);
        return sa ? sa->connectedGoalAreas : std::vector<int>();
    }
    std::vector<int> result;
    if (type == GOAL && GetGoalArea(sourceID)) {
        for (auto& kv : m_SpawnAreas)
            for (int g : kv.second.connectedGoalAreas)
                if (g == sourceID) { result.push_back(kv.first); break; }
    }
    return result;
}

void Simulator::FindNNearestNeighbors(const Entity& agent, int n, std::vector<Entity>& outNeighbors, int& outNNeighbors) {  // Simulator.cpp:211-227
    if ((int)outNeighbors.size() != n || n != 5) {
        printf("ERROR: FindNNearestNeighbors() expects std::vector<Entity>& outNeighbors to be of size n (= 5).\n");
        return;
    }
    if (!m_NeighborsValid) {  // one device query serves every agent until the state changes
        const int count = m_LastEntityIdx + 1;
        m_NeighborIds.assign(5 * (size_t)std::max(count, 1), -1);
        m_NeighborCounts.assign((size_t)std::max(count, 1), -1);
        Check(ecmgpu_find_neighbors(m_Gpu, count, m_NeighborIds.data(), m_NeighborCounts.data()), "ecmgpu_find_neighbors");
        m_NeighborsValid = true;
    }
    outNNeighbors = std::max(0, m_NeighborCounts[agent]);
    for (int k = 0; k < 5; k++) outNeighbors[k] = m_NeighborIds[5 * (size_t)agent + k];
    if (NN_TO_DRAW == agent) NEAREST_NEIGHBORS = outNeighbors;
}

// Simulator::FindNearestObstacles (Simulator.cpp:259-292) on the host mirror; a query API, not the hot path.
void Simulator::FindNNearestNeighborsDeprecated(const Entity& agent, int n, std::vector<Entity>& outNeighbors, int& outNNeighbors) {
    outNeighbors.clear();
    outNeighbors.resize(n);  // Simulator.cpp:230-231
    FindNNearestNeighbors(agent, n, outNeighbors, outNNeighbors);
}

void KDTree::KNearestAgents(Simulator* simulation, int agent, int k, std::vector<Entity>& outAgents, int& outNumNeighbors) {
    simulation->FindNNearestNeighbors(agent, k, outAgents, outNumNeighbors);
}

int Simulator::AddObstacleArea(const Point& position, const Vec2& halfSize, bool updateECM) {  // Simulator.cpp:383-412
    ObstacleArea oa;
    oa.ID = m_ObstacleAreas.size() == 0 ? 0 : m_ObstacleAreas[m_ObstacleAreas.size() - 1].ID + 1;
    oa.Position = position;
    oa.HalfHeight = halfSize.y;
    oa.HalfWidth = halfSize.x;
    oa.obstacleVerts.push_back(Point(position.x + halfSize.x, position.y + halfSize.y));
    oa.obstacleVerts.push_back(Point(position.x - halfSize.x, position.y + halfSize.y));
    oa.obstacleVerts.push_back(Point(position.x - halfSize.x, position.y - halfSize.y));
    oa.obstacleVerts.push_back(Point(position.x + halfSize.x, position.y - halfSize.y));
    // Environment::AddObstacle (Environment.cpp:198-229): the obstacle joins the list FindNearestObstacles scans
    float xy[8];
    for (int i = 0; i < 4; i++) { xy[2 * i] = oa.obstacleVerts[i].x; xy[2 * i + 1] = oa.obstacleVerts[i].y; }
    ecmb200::AppendObstacle(m_Obst, xy, 4);
    const auto& o = m_Obst;
    Check(ecmgpu_set_obstacles(m_Gpu, o.num_vertices(), o.xy.data(), o.next.data(), o.prev.data(), o.convex.data()), "ecmgpu_set_obstacles");
    if (updateECM) m_Error = "AddObstacleArea: updateECM needs the host-side ECM generator (not part of this library); the ECM and the planned paths are unchanged";
    return oa.ID;
}

void Simulator::FindNearestObstacles(const Entity& agent, float rangeSquared, std::vector<int>& outObstacles) const {
    const auto& o = m_Obst;
    const float ax = m_Positions[agent].x, ay = m_Positions[agent].y;
    auto approx = [](float px, float py, float qx, float qy) {
        const float E = 0.0001f;
        return px < (qx + E) && px > (qx - E) && py < (qy + E) && py > (qy - E);
    };
    for (int i = 0; i < o.num_vertices(); i++) {
        const float px = o.xy[2 * i], py = o.xy[2 * i + 1], qx = o.xy[2 * o.next[i]], qy = o.xy[2 * o.next[i] + 1];
        const float sl = (px - ax) * (qy - py) - (py - ay) * (qx - px);
        const float dx = qx - px, dy = qy - py;
        const float sq = (sl * sl) / (dx * dx + dy * dy);
        if (sq < rangeSquared && sl < 0.0f) {
            float cx = px, cy = py;
            if (!approx(px, py, qx, qy)) {
                float d = ((ax - px) * dx + (ay - py) * dy) / (dx * dx + dy * dy);
                if (d > 1.0f) d = 1.0f;
                if (d < 0.0f) d = 0.0f;
                cx = px + d * dx;
                cy = py + d * dy;
            }
            const float ex = ax - cx, ey = ay - cy;
            if (ex * ex + ey * ey < rangeSquared) outObstacles.push_back(i);
        }
    }
}

}  // namespace Simulation
}  // namespace ECM
