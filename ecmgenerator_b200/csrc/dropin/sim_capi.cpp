// C wrappers around the drop-in Simulator class so that the Python test-suite (ctypes) can drive it
// the way the reference's Application drives its Simulator (Application.cpp:137-144,
// ECMRenderer.cpp:836-884: Update once per frame, then read the raw component arrays).
#include <cstdint>
#include <cstring>
#include <exception>
#include <string>

#include "../../../include/ecm_b200.h"
#include "../../../include/ecm_b200_host.h"
#include "Simulator.h"

using ECM::Point;
using ECM::Vec2;
using ECM::Simulation::Simulator;

struct ecmhost_world {  // same definition as host_capi.cpp (one library, one layout)
    ecmb200::FlatWorld w;
};

namespace {
struct SimBox {
    ecmb200::PathPlanner* planner = nullptr;
    Simulator* sim = nullptr;
};
thread_local std::string g_err;
}  // namespace

extern "C" {

const char* ecmsim_last_error() { return g_err.c_str(); }

void* ecmsim_create(const ecmhost_world* w, int max_agents, float step, int device) {
    if (!w) { g_err = "null world"; return nullptr; }
    SimBox* b = new SimBox();
    try {
        b->planner = new ecmb200::PathPlanner(&w->w);
        b->sim = new Simulator(&w->w, b->planner, max_agents, step, device);
        b->sim->Initialize();
    } catch (const std::exception& e) {
        g_err = e.what();
        delete b->sim;
        delete b->planner;
        delete b;
        return nullptr;
    }
    return b;
}

void ecmsim_destroy(void* h) {
    SimBox* b = (SimBox*)h;
    if (!b) return;
    delete b->sim;
    delete b->planner;
    delete b;
}

#define GUARD(expr, fail_value)          \
    try { expr; }                        \
    catch (const std::exception& e) {    \
        g_err = e.what();                \
        return fail_value;               \
    }

int ecmsim_spawn_agent(void* h, float sx, float sy, float gx, float gy, float clearance, float speed) {
    GUARD(return ((SimBox*)h)->sim->SpawnAgent(Point(sx, sy), Point(gx, gy), clearance, speed), -2)
}
int ecmsim_destroy_agent(void* h, int idx) { GUARD(((SimBox*)h)->sim->DestroyAgent(idx); return 0, -2) }
int ecmsim_update(void* h, float dt) { GUARD(((SimBox*)h)->sim->Update(dt); return 0, -2) }
int ecmsim_reset(void* h) { GUARD(((SimBox*)h)->sim->Reset(); return 0, -2) }
int ecmsim_update_path(void* h, int e, float x, float y, float gx, float gy) {
    GUARD(((SimBox*)h)->sim->UpdatePath(e, Point(x, y), Point(gx, gy)); return 0, -2)
}
int ecmsim_add_position(void* h, int e, float x, float y) { GUARD(((SimBox*)h)->sim->AddPosition(e, x, y); return 0, -2) }
int ecmsim_num_agents(void* h) { return ((SimBox*)h)->sim->GetNumAgents(); }
int ecmsim_last_index(void* h) { return ((SimBox*)h)->sim->GetLastIndex(); }
const float* ecmsim_positions(void* h) { return (const float*)((SimBox*)h)->sim->GetPositionData(); }
const float* ecmsim_velocities(void* h) { return (const float*)((SimBox*)h)->sim->GetVelocityData(); }
const float* ecmsim_preferred_velocities(void* h) { return (const float*)((SimBox*)h)->sim->GetPreferredVelocityData(); }
const float* ecmsim_attraction_points(void* h) { return (const float*)((SimBox*)h)->sim->GetAttractionPointData(); }
const float* ecmsim_clearances(void* h) { return (const float*)((SimBox*)h)->sim->GetClearanceData(); }
const uint8_t* ecmsim_active_flags(void* h) {
    static_assert(sizeof(bool) == 1, "bool mirrors are read as bytes");
    return (const uint8_t*)((SimBox*)h)->sim->GetActiveFlags();
}
int ecmsim_path(void* h, int e, float* out_xy, int cap) {
    const auto& pc = ((SimBox*)h)->sim->GetPathData()[e];
    for (int j = 0; j < pc.numPoints && j < cap; j++) { out_xy[2 * j] = pc.x[j]; out_xy[2 * j + 1] = pc.y[j]; }
    return pc.numPoints;
}
int ecmsim_valid_spawn_location(void* h, float x, float y, float clearance) {
    return ((SimBox*)h)->sim->ValidSpawnLocation(Point(x, y), clearance) ? 1 : 0;
}
int ecmsim_find_neighbors(void* h, int agent, int* out5) {
    std::vector<int> nb(5, -1);
    int n = 0;
    GUARD(((SimBox*)h)->sim->FindNNearestNeighbors(agent, 5, nb, n), -2)
    for (int k = 0; k < 5; k++) out5[k] = nb[k];
    return n;
}
int ecmsim_find_obstacles(void* h, int agent, float range_squared, int* out, int cap) {
    std::vector<int> v;
    ((SimBox*)h)->sim->FindNearestObstacles(agent, range_squared, v);
    for (int k = 0; k < (int)v.size() && k < cap; k++) out[k] = v[k];
    return (int)v.size();
}
int ecmsim_add_spawn_area(void* h, float x, float y, float hw, float hh, float clearance, float speed) {
    ECM::Simulation::SpawnConfiguration cfg;
    cfg.clearanceMin = clearance;
    cfg.preferredSpeedMin = speed;
    return ((SimBox*)h)->sim->AddSpawnArea(Point(x, y), Vec2(hw, hh), cfg);
}
int ecmsim_add_goal_area(void* h, float x, float y, float hw, float hh) { return ((SimBox*)h)->sim->AddGoalArea(Point(x, y), Vec2(hw, hh)); }
void ecmsim_set_spawn_mode(void* h, int mode, unsigned long long seed) { ((SimBox*)h)->sim->SetSpawnMode((Simulator::SpawnMode)mode, seed); }
void ecmsim_spawn_checks(void* h, long long out[2]) { out[0] = ((SimBox*)h)->sim->SpawnChecksOnDevice(); out[1] = ((SimBox*)h)->sim->SpawnChecksOnHost(); }
void ecmsim_connect_areas(void* h, int spawn_id, int goal_id, float rate) { ((SimBox*)h)->sim->ConnectSpawnGoalAreas(spawn_id, goal_id, rate); }
int ecmsim_add_obstacle_area(void* h, float x, float y, float hw, float hh, int update_ecm) {
    GUARD(return ((SimBox*)h)->sim->AddObstacleArea(Point(x, y), Vec2(hw, hh), update_ecm != 0), -2)
}
int ecmsim_num_obstacle_vertices(void* h) { return ((SimBox*)h)->sim->GetObstacles().num_vertices(); }
// the reference's other routes to the neighbour query: GetKDTree()->KNearestAgents and the deprecated brute force
int ecmsim_find_neighbors_via(void* h, int agent, int route, int* out5) {
    std::vector<int> nb(5, -1);
    int n = 0;
    auto* sim = ((SimBox*)h)->sim;
    GUARD(if (route == 0) sim->GetKDTree()->KNearestAgents(sim, agent, 5, nb, n); else sim->FindNNearestNeighborsDeprecated(agent, 5, nb, n), -2)
    for (int k = 0; k < 5; k++) out5[k] = nb[k];
    return n;
}
// parity runs against the UNMODIFIED reference: the GPU tick with the reference's own KD-tree lists (ecm_b200.h)
int ecmsim_set_neighbor_mode(void* h, int mode) {
    auto* sim = ((SimBox*)h)->sim;
    const int rc = ecmgpu_set_neighbor_mode(sim->GetGpuHandle(), mode);
    if (rc) g_err = ecmgpu_last_error(sim->GetGpuHandle());
    return rc;
}

}  // extern "C"
