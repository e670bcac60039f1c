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
    : world_(world), planner_(planner), device_(device), capacity_(maxAgents), tick_seconds_(simStepTime) {
    for (int i = maxAgents - 1; i >= 0; i--) free_slots_.push(i);  // Simulator.h:66-69
}

Simulator::~Simulator() { ReleaseAll(); }

void Simulator::Check(int rc, const char* what) {
    if (rc == ECMGPU_OK) return;
    error_ = std::string(what) + ": " + (gpu_ ? ecmgpu_last_error(gpu_) : ecmgpu_last_error(nullptr));
    throw std::runtime_error(error_);  // no CPU fallback: a GPU failure is fatal for the simulation
}

void Simulator::Initialize() {  // Simulator.cpp:21-60
    const int n = capacity_;
    last_slot_ = -1;
    alive_ = new bool[n]();
    xy_ = new PositionComponent[n]();
    attraction_ = new PositionComponent[n]();
    vel_ = new VelocityComponent[n]();
    pref_vel_ = new VelocityComponent[n]();
    pref_speed_ = new SpeedComponent[n]();
    radius_ = new ClearanceComponent[n]();
    routes_ = new PathComponent[n];
    for (int i = 0; i < n; i++) {
        routes_[i].x = nullptr;
        routes_[i].y = nullptr;
        routes_[i].currentIndex = -1;
        routes_[i].numPoints = 0;
    }
    ecmgpu_params prm;
    memset(&prm, 0, sizeof(prm));
    prm.device = device_;
    prm.max_agents = n;
    prm.step = tick_seconds_;
    prm.record_neighbors = 0;
    Check(ecmgpu_create(&prm, &gpu_), "ecmgpu_create");
    const auto& e = world_->ecm;
    Check(ecmgpu_set_ecm(gpu_, world_->bbox, e.num_vertices(), e.vert_xy.data(), e.vert_clear.data(), e.num_edges(),
                         e.edge_v.data(), e.edge_cl.data()), "ecmgpu_set_ecm");
    // the half-edge rings let ecmgpu_plan_paths(GetGpuHandle(), ...) plan routes in batches on the device; the tick does
    // not need them, so a graph they do not describe only leaves that entry point unavailable
    if ((int)e.vert_he.size() == e.num_vertices() && (int)e.he_next.size() == 2 * e.num_edges())
        (void)ecmgpu_set_ecm_topology(gpu_, e.vert_he.data(), e.he_next.data());
    obstacles_ = world_->obst;
    const auto& o = obstacles_;
    Check(ecmgpu_set_obstacles(gpu_, o.num_vertices(), o.xy.data(), o.next.data(), o.prev.data(), o.convex.data()), "ecmgpu_set_obstacles");
    printf("SIMULATOR: Data for %d agents was created.\n", n);
}

void Simulator::ReleaseAll() {  // Simulator.cpp:62-95
    if (routes_) {
        for (int i = 0; i < capacity_; i++) {
            delete[] routes_[i].x;
            delete[] routes_[i].y;
        }
    }
    delete[] xy_; delete[] attraction_; delete[] vel_; delete[] pref_vel_;
    delete[] pref_speed_; delete[] routes_; delete[] alive_; delete[] radius_;
    xy_ = attraction_ = nullptr;
    vel_ = pref_vel_ = nullptr;
    pref_speed_ = nullptr; routes_ = nullptr; alive_ = nullptr; radius_ = nullptr;
    if (gpu_) {
        ecmgpu_destroy(gpu_);
        gpu_ = nullptr;
        printf("SIMULATOR: Data was destroyed.\n");
    }
}

void Simulator::StoreRoute(int e, const std::vector<ecmb200::P2f>& path) {  // Simulator.cpp:114-123
    PathComponent& pc = routes_[e];
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
    Check(ecmgpu_set_path(gpu_, e, xy.data(), n), "ecmgpu_set_path");
}

// Simulator::UpdatePath (Simulator.cpp:97-124).  A failed query keeps the previous path.
void Simulator::UpdatePath(const Entity& e, const Point& location, const Point& goal) {
    std::vector<ecmb200::P2f> path;
    const bool ok = planner_->FindPath(ecmb200::P2f{location.x, location.y}, ecmb200::P2f{goal.x, goal.y}, radius_[e].clearance, path);
    if (!ok || path.size() < 2) {
        if (routes_[e].numPoints >= 2) {  // re-arm the replan request instead of storing an unusable path
            std::vector<ecmb200::P2f> old((size_t)routes_[e].numPoints);
            for (int j = 0; j < routes_[e].numPoints; j++) old[j] = ecmb200::P2f{routes_[e].x[j], routes_[e].y[j]};
            StoreRoute(e, old);
        }
        return;
    }
    StoreRoute(e, path);
}

int Simulator::SpawnAgent(const Point& start, const Point& goal, float clearance, float preferredSpeed) {  // Simulator.cpp:168-200
    if (free_slots_.empty()) return -1;
    if (!ValidSpawnLocation(start, clearance)) return -1;
    std::vector<ecmb200::P2f> path;
    if (!planner_->FindPath(ecmb200::P2f{start.x, start.y}, ecmb200::P2f{goal.x, goal.y}, clearance, path) || path.size() < 2) return -1;
    count_++;
    int idx = free_slots_.top();
    free_slots_.pop();
    last_slot_ = last_slot_ < idx ? idx : last_slot_;
    xy_[idx].x = start.x;
    xy_[idx].y = start.y;
    radius_[idx].clearance = clearance;
    pref_speed_[idx].speed = preferredSpeed;
    alive_[idx] = true;
    std::vector<float> xy(2 * path.size());
    for (size_t j = 0; j < path.size(); j++) { xy[2 * j] = path[j].x; xy[2 * j + 1] = path[j].y; }
    Check(ecmgpu_spawn(gpu_, idx, start.x, start.y, clearance, preferredSpeed, xy.data(), (int)path.size()), "ecmgpu_spawn");
    PathComponent& pc = routes_[idx];
    delete[] pc.x;
    delete[] pc.y;
    pc.numPoints = (int)path.size();
    pc.currentIndex = 0;
    pc.x = new float[path.size()];
    pc.y = new float[path.size()];
    for (size_t j = 0; j < path.size(); j++) { pc.x[j] = path[j].x; pc.y[j] = path[j].y; }
    pref_vel_[idx].dx = pref_vel_[idx].dy = 0.0f;
    vel_[idx].dx = vel_[idx].dy = 0.0f;
    attraction_[idx].x = attraction_[idx].y = 0.0f;
    nbr_valid_ = false;
    return idx;
}

void Simulator::DestroyAgent(int idx) {  // Simulator.cpp:202-208
    count_--;
    alive_[idx] = false;
    free_slots_.push(idx);
    Check(ecmgpu_destroy_agent(gpu_, idx), "ecmgpu_destroy_agent");
    nbr_valid_ = false;
}

void Simulator::AddPosition(Entity entity, float x, float y) {  // Simulator.h:87-90
    xy_[entity].x = x;
    xy_[entity].y = y;
    Check(ecmgpu_write(gpu_, ECMGPU_POS, &xy_[entity], entity, 1), "ecmgpu_write");
    nbr_valid_ = false;
}

bool Simulator::ValidSpawnLocation(const Point& location, float clearance) const {  // Simulator.cpp:295-311
    float clearanceSquared = clearance * clearance;
    for (int i = 0; i <= last_slot_; i++) {
        if (!alive_[i]) continue;
        float dx = location.x - xy_[i].x, dy = location.y - xy_[i].y;
        float d = dx * dx + dy * dy;
        if (d < clearanceSquared) return false;
    }
    return true;
}

void Simulator::TrimLastSlot() {  // Simulator.cpp:481-492
    int emptyCounter = 0;
    for (int i = last_slot_; i >= 0; i--) {
        if (alive_[i]) break;
        emptyCounter++;
    }
    last_slot_ = last_slot_ - emptyCounter;
}

void Simulator::RunSpawnAreas() {  // Simulator.cpp:494-536
    for (auto iter = spawn_areas_.begin(); iter != spawn_areas_.end(); iter++) {
        SpawnArea& area = iter->second;
        for (int ga = 0; ga < (int)area.connectedGoalAreas.size(); ga++) {
            area.timeSinceLastSpawn[ga] += tick_seconds_;
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
                Point goal = goal_areas_[area.connectedGoalAreas[ga]].GetRandomPositionInArea();
                SpawnAgent(start, goal, clearance, speed);
            }
            area.timeSinceLastSpawn[ga] -= (float)agentsToSpawn / area.spawnRate[ga];
        }
    }
}

void Simulator::Update(float /*dt*/) {  // Simulator.cpp:314-323
    TrimLastSlot();
    RunSpawnAreas();
    nbr_valid_ = false;
    if (last_slot_ < 0) return;
    Check(ecmgpu_update(gpu_), "ecmgpu_update");
    // events of this tick: arrivals (Simulator.cpp:564-566) then replans (Simulator.cpp:581-587), in slot
    // order like the reference's loop; the mirrors still hold the PRE-tick positions the replans start from
    const int count = last_slot_ + 1;
    event_scratch_.resize(2 * (size_t)count);
    int nr = 0, nd = 0;
    Check(ecmgpu_poll_events(gpu_, event_scratch_.data(), count, &nr, event_scratch_.data() + count, count, &nd), "ecmgpu_poll_events");
    for (int k = 0; k < nd; k++) {
        const int e = event_scratch_[count + k];
        count_--;
        alive_[e] = false;
        free_slots_.push(e);
        // the reference assigns the goal as attraction point before destroying (Simulator.cpp:559-562)
        attraction_[e].x = routes_[e].x[routes_[e].numPoints - 1];
        attraction_[e].y = routes_[e].y[routes_[e].numPoints - 1];
    }
    for (int k = 0; k < nr; k++) {
        const int e = event_scratch_[k];
        printf("Recalculate path...\n");
        Point cur(xy_[e].x, xy_[e].y);
        Point goal(routes_[e].x[routes_[e].numPoints - 1], routes_[e].y[routes_[e].numPoints - 1]);
        UpdatePath(e, cur, goal);
    }
    // refresh the mirrors the getters expose; destroyed agents keep their stale components (Appendix B.5)
    std::vector<PositionComponent> pos(count), att(count);
    std::vector<VelocityComponent> vel(count), pref(count);
    Check(ecmgpu_read(gpu_, ECMGPU_POS, pos.data(), 0, count), "ecmgpu_read");
    Check(ecmgpu_read(gpu_, ECMGPU_VEL, vel.data(), 0, count), "ecmgpu_read");
    Check(ecmgpu_read(gpu_, ECMGPU_PREFVEL, pref.data(), 0, count), "ecmgpu_read");
    Check(ecmgpu_read(gpu_, ECMGPU_ATTRACTION, att.data(), 0, count), "ecmgpu_read");
    for (int i = 0; i < count; i++) {
        if (!alive_[i]) continue;
        xy_[i] = pos[i];
        vel_[i] = vel[i];
        pref_vel_[i] = pref[i];
        attraction_[i] = att[i];
    }
}

void Simulator::Reset() {  // Simulator.cpp:325-333
    for (int i = 0; i <= last_slot_; i++) {
        if (!alive_[i]) continue;
        DestroyAgent(i);
    }
}

int Simulator::AddSpawnArea(const Point& position, const Vec2& halfSize, const SpawnConfiguration& config, int ID) {  // Simulator.cpp:336-360
    SpawnArea sa;
    sa.HalfWidth = halfSize.x;
    sa.HalfHeight = halfSize.y;
    if (ID == -1) { sa.ID = next_spawn_id_; next_spawn_id_++; }
    else sa.ID = ID;
    sa.Position = position;
    sa.spawnConfiguration = config;
    spawn_areas_.emplace(sa.ID, sa);
    return sa.ID;
}

int Simulator::AddGoalArea(const Point& position, const Vec2& halfSize, int ID) {  // Simulator.cpp:362-381
    GoalArea ga;
    if (ID == -1) { ga.ID = next_goal_id_; next_goal_id_++; }
    else ga.ID = ID;
    ga.Position = position;
    ga.HalfHeight = halfSize.y;
    ga.HalfWidth = halfSize.x;
    goal_areas_.emplace(ga.ID, ga);
    return ga.ID;
}

void Simulator::RemoveArea(SimAreaType areaType, int ID) {  // Simulator.cpp:416-426
    if (areaType == SPAWN) spawn_areas_.erase(ID);
    if (areaType == GOAL) goal_areas_.erase(ID);
}

void Simulator::ConnectSpawnGoalAreas(int spawnID, int goalID, float spawnRate) {  // Simulator.cpp:428-440
    for (int ga : spawn_areas_[spawnID].connectedGoalAreas)
        if (ga == goalID) return;
    SpawnArea& sa = spawn_areas_[spawnID];
    sa.connectedGoalAreas.push_back(goalID);
    sa.timeSinceLastSpawn.push_back(0.0f);
    sa.spawnRate.push_back(spawnRate);
}

void Simulator::DeconnectSpawnGoalAreas(int spawnID, int goalID) {  // Simulator.cpp:442-460
    if (spawnID < 0 || spawnID >= (int)spawn_areas_.size()) return;
    auto& v = spawn_areas_[spawnID].connectedGoalAreas;
    for (size_t i = 0; i < v.size(); i++)
        if (v[i] == goalID) { v.erase(v.begin() + i); return; }
}

SpawnArea* Simulator::GetSpawnArea(int ID) {
    auto it = spawn_areas_.find(ID);
    return it != spawn_areas_.end() ? &it->second : nullptr;
}
GoalArea* Simulator::GetGoalArea(int ID) {
    auto it = goal_areas_.find(ID);
    return it != goal_areas_.end() ? &it->second : nullptr;
}

std::vector<int> Simulator::GetConnectedAreas(int sourceID, SimAreaType type) {  // Simulator.cpp:126-166
    if (type == SPAWN) {
        SpawnArea* sa = GetSpawnArea(sourceID);
        return sa ? sa->connectedGoalAreas : std::vector<int>();
    }
    std::vector<int> result;
    if (type == GOAL && GetGoalArea(sourceID)) {
        for (auto& kv : spawn_areas_)
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
    if (!nbr_valid_) {  // one device query serves every agent until the state changes
        const int count = last_slot_ + 1;
        nbr_ids_.assign(5 * (size_t)std::max(count, 1), -1);
        nbr_counts_.assign((size_t)std::max(count, 1), -1);
        Check(ecmgpu_find_neighbors(gpu_, count, nbr_ids_.data(), nbr_counts_.data()), "ecmgpu_find_neighbors");
        nbr_valid_ = true;
    }
    outNNeighbors = std::max(0, nbr_counts_[agent]);
    for (int k = 0; k < 5; k++) outNeighbors[k] = nbr_ids_[5 * (size_t)agent + k];
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
    oa.ID = obstacle_areas_.size() == 0 ? 0 : obstacle_areas_[obstacle_areas_.size() - 1].ID + 1;
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
    ecmb200::AppendObstacle(obstacles_, xy, 4);
    const auto& o = obstacles_;
    Check(ecmgpu_set_obstacles(gpu_, o.num_vertices(), o.xy.data(), o.next.data(), o.prev.data(), o.convex.data()), "ecmgpu_set_obstacles");
    if (updateECM) error_ = "AddObstacleArea: updateECM needs the host-side ECM generator (not part of this library); the ECM and the planned paths are unchanged";
    return oa.ID;
}

void Simulator::FindNearestObstacles(const Entity& agent, float rangeSquared, std::vector<int>& outObstacles) const {
    const auto& o = obstacles_;
    const float ax = xy_[agent].x, ay = xy_[agent].y;
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
