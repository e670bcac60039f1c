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
    return SpawnChecked(start, goal, clearance, preferredSpeed);
}

int Simulator::SpawnChecked(const Point& start, const Point& goal, float clearance, float preferredSpeed) {
    if (free_slots_.empty()) return -1;
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
    spawned_this_tick_.push_back(start);
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

// ---- rand() checkpoints ------------------------------------------------------------------------------------------
// glibc's rand() is random() on a state array that setstate(3) can switch: swapping to a scratch state hands out the address
// of the live one and writes its read position into its first word, so a memcpy of the array IS the generator.
namespace {
#if defined(__GLIBC__)
struct RandCheckpoint {
    unsigned char data[256];
    size_t bytes = 0;
};
char* ScratchRandState() {
    static char scratch[128];
    static bool ready = false;
    if (!ready) {
        char* live = initstate(1u, scratch, sizeof(scratch));  // initialises `scratch` and switches to it ...
        setstate(live);                                        // ... so switch straight back
        ready = true;
    }
    return scratch;
}
size_t RandStateBytes(const char* state) {
    int32_t word0;
    memcpy(&word0, state, 4);
    static const size_t kBytes[5] = {8, 32, 64, 128, 256};  // TYPE_0 .. TYPE_4 (stdlib/random_r.c)
    return kBytes[((word0 % 5) + 5) % 5];
}
bool SaveRand(RandCheckpoint& cp) {
    char* live = setstate(ScratchRandState());
    if (!live) return false;
    cp.bytes = RandStateBytes(live);
    memcpy(cp.data, live, cp.bytes);
    setstate(live);
    return true;
}
void RestoreRand(const RandCheckpoint& cp) {
    char* live = setstate(ScratchRandState());
    memcpy(live, cp.data, cp.bytes);
    setstate(live);
}
constexpr bool kRandCheckpoints = true;
#else
struct RandCheckpoint {};
bool SaveRand(RandCheckpoint&) { return false; }
void RestoreRand(const RandCheckpoint&) {}
constexpr bool kRandCheckpoints = false;
#endif
}  // namespace

bool Simulator::ClashesWithThisTick(const Point& p, float clearance) const {
    const float c2 = clearance * clearance;
    for (const Point& q : spawned_this_tick_) {
        const float dx = p.x - q.x, dy = p.y - q.y;
        if (dx * dx + dy * dy < c2) return true;  // the expression of ValidSpawnLocation (Simulator.cpp:303-306)
    }
    return false;
}

// The reference's loop, attempt by attempt on the host mirrors (Simulator.cpp:494-536).
void Simulator::RunSpawnAreasSequential() {
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
                    spawn_checks_host_++;
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

void Simulator::RunSpawnAreas() {  // Simulator.cpp:494-536
    spawned_this_tick_.clear();
    if (spawn_mode_ == SPAWN_RAND_SEQUENTIAL || (spawn_mode_ == SPAWN_RAND_BATCHED && !kRandCheckpoints)) {
        RunSpawnAreasSequential();
        return;
    }
    // The requests of a tick - which area, towards which goal area, how many - do not depend on any random number or
    // validity test: only the DRAWS do.  Collect them in the reference's order, then answer them in batches.
    std::vector<SpawnRequest> req;
    for (auto iter = spawn_areas_.begin(); iter != spawn_areas_.end(); iter++) {
        SpawnArea& area = iter->second;
        for (int ga = 0; ga < (int)area.connectedGoalAreas.size(); ga++) {
            area.timeSinceLastSpawn[ga] += tick_seconds_;
            int agentsToSpawn = area.timeSinceLastSpawn[ga] * area.spawnRate[ga];
            for (int i = 0; i < agentsToSpawn; i++)
                req.push_back(SpawnRequest{&area, area.connectedGoalAreas[ga], area.spawnConfiguration.clearanceMin, area.spawnConfiguration.preferredSpeedMin});
            area.timeSinceLastSpawn[ga] -= (float)agentsToSpawn / area.spawnRate[ga];
        }
    }
    if (req.empty()) return;
    if (spawn_mode_ == SPAWN_DEVICE_COUNTER) RunSpawnRequestsOnDevice(req);
    else RunSpawnRequestsBatched(req);
}

// rand() in the reference's order, validity on the GPU.  Speculation: every request succeeds at its first attempt, so the
// stream is start(3 numbers), goal(3), start, goal, ...; all starts of the round are tested in ONE device call.  The
// first request whose start turns out invalid stops the round: the generator goes back to the state right after that
// start was drawn and the request is finished attempt by attempt (attempts 2 .. 10, one device test each), then the
// next round begins.  Agents spawned earlier in the tick are not on the device yet when a batch is tested: the few of
// them are compared on the host, with ValidSpawnLocation's own expression.
void Simulator::RunSpawnRequestsBatched(const std::vector<SpawnRequest>& req) {
    const int maxSpawnAttempts = 10;
    size_t next = 0;
    std::vector<Point> start, goal;
    std::vector<RandCheckpoint> after_start;
    std::vector<float> xy, cl;
    std::vector<uint8_t> ok;
    while (next < req.size()) {
        const size_t m = req.size() - next;
        start.resize(m); goal.resize(m); after_start.resize(m); xy.resize(2 * m); cl.resize(m); ok.assign(m, 0);
        for (size_t k = 0; k < m; k++) {
            const SpawnRequest& r = req[next + k];
            start[k] = r.area->GetRandomPositionInArea();
            SaveRand(after_start[k]);
            goal[k] = goal_areas_[r.goalArea].GetRandomPositionInArea();
            xy[2 * k] = start[k].x; xy[2 * k + 1] = start[k].y;
            cl[k] = r.clearance;
        }
        Check(ecmgpu_valid_spawn_locations(gpu_, (int)m, xy.data(), cl.data(), ok.data()), "ecmgpu_valid_spawn_locations");
        spawn_checks_device_ += (long long)m;
        size_t k = 0;
        for (; k < m; k++) {
            const SpawnRequest& r = req[next + k];
            if (!ok[k] || ClashesWithThisTick(start[k], r.clearance)) break;
            SpawnChecked(start[k], goal[k], r.clearance, r.speed);
        }
        next += k;
        if (k == m) break;
        // request `next` failed its first attempt: everything drawn after its start is undone
        const SpawnRequest& r = req[next];
        RestoreRand(after_start[k]);
        bool found = false;
        Point s2;
        for (int attempt = 1; attempt < maxSpawnAttempts && !found; attempt++) {
            s2 = r.area->GetRandomPositionInArea();
            const float p[2] = {s2.x, s2.y};
            uint8_t v = 0;
            Check(ecmgpu_valid_spawn_locations(gpu_, 1, p, &r.clearance, &v), "ecmgpu_valid_spawn_locations");
            spawn_checks_device_++;
            found = v != 0 && !ClashesWithThisTick(s2, r.clearance);
        }
        if (!found) {
            printf("Could not find a valid spawn position in the spawn area!\n");
        } else {
            Point g2 = goal_areas_[r.goalArea].GetRandomPositionInArea();
            SpawnChecked(s2, g2, r.clearance, r.speed);
        }
        next++;
    }
}

// Draws from the counter-based generator on the device (ecmgpu_draw_spawns): one call per tick.
void Simulator::RunSpawnRequestsOnDevice(const std::vector<SpawnRequest>& req) {
    const size_t m = req.size();
    std::vector<float> sb(4 * m), gb(4 * m), cl(m), s_xy(2 * m), g_xy(2 * m);
    std::vector<uint8_t> ok(m, 0);
    for (size_t k = 0; k < m; k++) {
        const Area& a = *req[k].area;
        const Area& g = goal_areas_[req[k].goalArea];
        sb[4 * k] = a.Position.x - a.HalfWidth; sb[4 * k + 1] = a.Position.y - a.HalfHeight;
        sb[4 * k + 2] = a.Position.x + a.HalfWidth; sb[4 * k + 3] = a.Position.y + a.HalfHeight;
        gb[4 * k] = g.Position.x - g.HalfWidth; gb[4 * k + 1] = g.Position.y - g.HalfHeight;
        gb[4 * k + 2] = g.Position.x + g.HalfWidth; gb[4 * k + 3] = g.Position.y + g.HalfHeight;
        cl[k] = req[k].clearance;
    }
    Check(ecmgpu_draw_spawns(gpu_, (int)m, sb.data(), gb.data(), cl.data(), spawn_seed_, spawn_counter_++, 10, s_xy.data(), g_xy.data(), ok.data()),
          "ecmgpu_draw_spawns");
    spawn_checks_device_ += (long long)m;
    for (size_t k = 0; k < m; k++) {
        const Point s(s_xy[2 * k], s_xy[2 * k + 1]), g(g_xy[2 * k], g_xy[2 * k + 1]);
        if (!ok[k] || ClashesWithThisTick(s, req[k].clearance)) continue;  // dropped, like a request that found no valid position
        SpawnChecked(s, g, req[k].clearance, req[k].speed);
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
