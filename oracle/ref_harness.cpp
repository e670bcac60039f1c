// TEST INFRASTRUCTURE - NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load the library built from this file.
//
// C-callable harness around the UNMODIFIED reference hot path.  The reference translation units
// are compiled where they lie under /root/reference (see oracle/Makefile); this file adds
//   (1) a definition of ECMGenerator::GenerateECM (declared ECMGenerator/ECMGenerator.h:19) that
//       fills the ECMGraph from flat arrays through the `friend class ECMGenerator` seam
//       (ECM.h:54) - the reference's own generator needs Boost.Polygon, which is absent;
//   (2) bulk loading of agents that bypasses the O(N) ValidSpawnLocation scan per spawn
//       (Simulator.cpp:295-311) but otherwise performs SpawnAgent's writes (Simulator.cpp:175-199);
//   (3) read-back of private component arrays (m_Forces has no getter).
// Nothing here re-implements per-tick arithmetic: stepping calls Simulator::Update.
#define private public
#define protected public
#include "Simulator.h"
#include "ECM.h"
#include "ECMCellCollection.h"
#include "ECMGenerator.h"
#include "ECMPathPlanner.h"
#include "Environment.h"
#include "KDTree.h"
#include "ORCA.h"
#include "IRMPathFollower.h"
#include "UtilityFunctions.h"
#undef private
#undef protected

#include <cstring>

namespace {
struct FlatEcmInput {
    int nV = 0, nE = 0;
    const float* vert_xy = nullptr;
    const float* vert_clear = nullptr;
    const int* vert_he = nullptr;
    const int* edge_v = nullptr;
    const float* edge_cl = nullptr;
    const int* he_next = nullptr;
};
FlatEcmInput g_input;  // consumed by GenerateECM below
}  // namespace

namespace ECM {

ECM* ECMGenerator::GenerateECM(const Environment& /*environment*/, ECM* ecm) {
    if (ecm == nullptr) ecm = new ECM();
    ECMGraph& g = ecm->GetECMGraph();
    const FlatEcmInput& in = g_input;
    for (int v = 0; v < in.nV; v++) {
        g.AddVertex(Point(in.vert_xy[2 * v], in.vert_xy[2 * v + 1]));
    }
    for (int v = 0; v < in.nV; v++) {
        g.SetVertexClearance(v, in.vert_clear[v]);
        g.SetVertexHalfEdge(v, in.vert_he[v]);
    }
    for (int e = 0; e < in.nE; e++) {
        ECMEdge* edge = g.AddEdge();
        const float* c = in.edge_cl + 8 * e;
        const Point L0(c[0], c[1]), R0(c[2], c[3]), L1(c[4], c[5]), R1(c[6], c[7]);
        // ECMGenerator.cpp:165-168: he[0] -> v1 with (closestLeft1, closestRight1),
        //                           he[1] -> v0 with (closestRight2, closestLeft2)
        g.AddHalfEdge(edge->idx, in.edge_v[2 * e + 1], L0, R0, 0);
        g.AddHalfEdge(edge->idx, in.edge_v[2 * e], R1, L1, 1);
    }
    for (int e = 0; e < in.nE; e++) {
        g.SetNextEdge(e, 0, in.he_next[2 * e]);
        g.SetNextEdge(e, 1, in.he_next[2 * e + 1]);
    }
    g.ConstructECMCells();
    return ecm;
}

}  // namespace ECM

using namespace ECM;
using namespace ECM::Simulation;

struct RefSim {
    Environment env;
    ECM::ECM* ecm = nullptr;
    PathPlanning::ECMPathPlanner* planner = nullptr;
    Simulator* sim = nullptr;
    std::vector<Entity> nn_cache;  // plays the role of ORCA::m_NeighborCache in query_neighbors
};

static void set_path(Simulator* s, int idx, const float* xy, int n) {
    PathComponent& pc = s->m_Paths[idx];
    if (pc.numPoints > 0) {
        delete[] pc.x;
        delete[] pc.y;
    }
    pc.x = new float[n > 0 ? n : 1];
    pc.y = new float[n > 0 ? n : 1];
    pc.numPoints = n;
    pc.currentIndex = 0;
    for (int j = 0; j < n; j++) {
        pc.x[j] = xy[2 * j];
        pc.y[j] = xy[2 * j + 1];
    }
}

extern "C" {

void* ecmref_create(const float* bbox, int nV, const float* vert_xy, const float* vert_clear, const int* vert_he,
                    int nE, const int* edge_v, const float* edge_cl, const int* he_next, int nObst,
                    const int* obst_first, const float* obst_xy, int max_agents, float step) {
    RefSim* r = new RefSim();
    std::vector<Segment> wa;
    wa.push_back(Segment(bbox[0], bbox[1], bbox[2], bbox[1]));
    wa.push_back(Segment(bbox[2], bbox[1], bbox[2], bbox[3]));
    wa.push_back(Segment(bbox[0], bbox[3], bbox[2], bbox[3]));
    wa.push_back(Segment(bbox[0], bbox[3], bbox[0], bbox[1]));
    r->env.AddWalkableArea(wa);
    for (int o = 0; o < nObst; o++) {
        std::vector<Point> pts;
        for (int k = obst_first[o]; k < obst_first[o + 1]; k++) pts.push_back(Point(obst_xy[2 * k], obst_xy[2 * k + 1]));
        r->env.AddObstacle(pts);  // Environment.cpp:199-229
    }
    r->env.m_Dirty = false;
    g_input = FlatEcmInput{nV, nE, vert_xy, vert_clear, vert_he, edge_v, edge_cl, he_next};
    r->env.ComputeECM();  // -> our ECMGenerator::GenerateECM above
    g_input = FlatEcmInput{};
    r->ecm = r->env.GetECM();
    r->planner = new PathPlanning::ECMPathPlanner(&r->ecm->GetECMGraph());
    r->sim = new Simulator(r->ecm, r->planner, &r->env, max_agents, step);
    r->sim->Initialize();
    for (int i = 0; i < max_agents; i++) r->sim->m_ActiveAgents[i] = false;  // new bool[] is uninitialised
    r->nn_cache.resize(5);
    return r;
}

void ecmref_destroy(void* h) {
    RefSim* r = (RefSim*)h;
    delete r->sim;
    delete r->planner;
    delete r;
}

// Obstacle topology as the reference's Obstacle::Initialize computed it (ECMDataTypes.cpp:23-61).
int ecmref_get_obstacles(void* h, float* xy, int* next, int* prev, uint8_t* convex) {
    RefSim* r = (RefSim*)h;
    const auto& obs = r->env.GetObstacles();
    std::unordered_map<const ObstacleVertex*, int> id;
    int n = 0;
    for (const Obstacle& o : obs)
        for (const ObstacleVertex* v : o.verts) id[v] = n++;
    if (!xy) return n;
    int k = 0;
    for (const Obstacle& o : obs)
        for (const ObstacleVertex* v : o.verts) {
            xy[2 * k] = v->p.x;
            xy[2 * k + 1] = v->p.y;
            next[k] = id[v->nextObstacle];
            prev[k] = id[v->prevObstacle];
            convex[k] = v->isConvex ? 1 : 0;
            k++;
        }
    return n;
}

// ECMPathPlanner::FindPath with the arguments Simulator::UpdatePath passes (Simulator.cpp:108-112).
int ecmref_plan_path(void* h, float sx, float sy, float gx, float gy, float clearance, float* out_xy, int cap) {
    RefSim* r = (RefSim*)h;
    PathPlanning::Corridor dummy;
    std::vector<Segment> portal;
    PathPlanning::Path path;
    bool ok = r->planner->FindPath(r->env, Point(sx, sy), Point(gx, gy), clearance, 0.0f, dummy, portal, path);
    int n = (int)path.size();
    for (int i = 0; i < n && i < cap; i++) {
        out_xy[2 * i] = path[i].x;
        out_xy[2 * i + 1] = path[i].y;
    }
    return ok ? n : -1;
}

// SpawnAgent's effects (Simulator.cpp:175-199) without ValidSpawnLocation; path either planned by
// the reference planner (path_off == NULL) or given (path_off[i]..path_off[i+1] points in path_xy).
// Returns the number of agents whose path has >= 2 points; agents with an unusable path are skipped
// (the reference would index path.x[-1], Simulator.cpp:554) and get slot -1 in out_slots.
int ecmref_bulk_load(void* h, int n, const float* pos_xy, const float* goal_xy, const float* radius,
                     const float* speed, const int* path_off, const float* path_xy, int* out_slots) {
    RefSim* r = (RefSim*)h;
    Simulator* s = r->sim;
    int loaded = 0;
    for (int i = 0; i < n; i++) {
        if (out_slots) out_slots[i] = -1;
        if (s->m_freeEntitySpaces.empty()) break;
        std::vector<float> planned;
        const float* pxy;
        int np;
        if (path_off) {
            np = path_off[i + 1] - path_off[i];
            pxy = path_xy + 2 * path_off[i];
        } else {
            PathPlanning::Corridor dummy;
            std::vector<Segment> portal;
            PathPlanning::Path path;
            r->planner->FindPath(r->env, Point(pos_xy[2 * i], pos_xy[2 * i + 1]), Point(goal_xy[2 * i], goal_xy[2 * i + 1]),
                                 radius[i], 0.0f, dummy, portal, path);
            np = (int)path.size();
            for (const Point& p : path) {
                planned.push_back(p.x);
                planned.push_back(p.y);
            }
            pxy = planned.data();
        }
        if (np < 2) continue;
        s->m_NumEntities++;
        int idx = s->m_freeEntitySpaces.top();
        s->m_freeEntitySpaces.pop();
        s->m_LastEntityIdx = s->m_LastEntityIdx < idx ? idx : s->m_LastEntityIdx;
        s->m_Positions[idx].x = pos_xy[2 * i];
        s->m_Positions[idx].y = pos_xy[2 * i + 1];
        s->m_Clearances[idx].clearance = radius[i];
        s->m_PreferredSpeed[idx].speed = speed[i];
        s->m_ActiveAgents[idx] = true;
        set_path(s, idx, pxy, np);
        s->m_PreferredVelocities[idx].dx = s->m_PreferredVelocities[idx].dy = 0.0f;
        s->m_Velocities[idx].dx = s->m_Velocities[idx].dy = 0.0f;
        s->m_Forces[idx].dx = s->m_Forces[idx].dy = 0.0f;
        s->m_AttractionPoints[idx].x = s->m_AttractionPoints[idx].y = 0.0f;  // new[] leaves it uninitialised
        if (out_slots) out_slots[i] = idx;
        loaded++;
    }
    return loaded;
}

int ecmref_spawn(void* h, float sx, float sy, float gx, float gy, float clearance, float speed) {
    return ((RefSim*)h)->sim->SpawnAgent(Point(sx, sy), Point(gx, gy), clearance, speed);
}

// Area API as the editor uses it (Command.cpp:33-159): Simulator.cpp:336-381, :428-440.
int ecmref_add_spawn_area(void* h, float x, float y, float hw, float hh, float clearance, float speed) {
    SpawnConfiguration cfg;
    cfg.clearanceMin = clearance;
    cfg.preferredSpeedMin = speed;
    return ((RefSim*)h)->sim->AddSpawnArea(Point(x, y), Vec2(hw, hh), cfg);
}
int ecmref_add_goal_area(void* h, float x, float y, float hw, float hh) { return ((RefSim*)h)->sim->AddGoalArea(Point(x, y), Vec2(hw, hh)); }
void ecmref_connect_areas(void* h, int spawn_id, int goal_id, float rate) { ((RefSim*)h)->sim->ConnectSpawnGoalAreas(spawn_id, goal_id, rate); }
// Simulator::AddObstacleArea (Simulator.cpp:383-412) -> Environment::AddObstacle (Environment.cpp:198-229), updateECM = false
int ecmref_add_obstacle_area(void* h, float x, float y, float hw, float hh) { return ((RefSim*)h)->sim->AddObstacleArea(Point(x, y), Vec2(hw, hh), false); }
void ecmref_srand(unsigned seed) { srand(seed); }
int ecmref_valid_spawn_location(void* h, float x, float y, float c) { return ((RefSim*)h)->sim->ValidSpawnLocation(Point(x, y), c) ? 1 : 0; }

void ecmref_set_kinematics(void* h, int slot, float x, float y, float vx, float vy) {
    Simulator* s = ((RefSim*)h)->sim;
    s->m_Positions[slot].x = x;
    s->m_Positions[slot].y = y;
    s->m_Velocities[slot].dx = vx;
    s->m_Velocities[slot].dy = vy;
}

void ecmref_set_attraction(void* h, int slot, float x, float y) {
    Simulator* s = ((RefSim*)h)->sim;
    s->m_AttractionPoints[slot].x = x;
    s->m_AttractionPoints[slot].y = y;
}

void ecmref_destroy_agent(void* h, int slot) { ((RefSim*)h)->sim->DestroyAgent(slot); }  // Simulator.cpp:202-208

int ecmref_path_len(void* h, int slot) { return ((RefSim*)h)->sim->m_Paths[slot].numPoints; }

int ecmref_get_path(void* h, int slot, float* out_xy, int cap) {
    const PathComponent& pc = ((RefSim*)h)->sim->m_Paths[slot];
    for (int j = 0; j < pc.numPoints && j < cap; j++) {
        out_xy[2 * j] = pc.x[j];
        out_xy[2 * j + 1] = pc.y[j];
    }
    return pc.numPoints;
}

void ecmref_step(void* h, int nsteps) {
    Simulator* s = ((RefSim*)h)->sim;
    for (int i = 0; i < nsteps; i++) s->Update(s->GetSimulationStepTime());
}

int ecmref_num_agents(void* h) { return ((RefSim*)h)->sim->GetNumAgents(); }
int ecmref_last_index(void* h) { return ((RefSim*)h)->sim->GetLastIndex(); }

// Copies slots [0, count): every array may be NULL.
void ecmref_get_state(void* h, int count, float* pos, float* vel, float* prefvel, float* attraction, float* force,
                      uint8_t* active) {
    Simulator* s = ((RefSim*)h)->sim;
    for (int i = 0; i < count; i++) {
        if (pos) { pos[2 * i] = s->m_Positions[i].x; pos[2 * i + 1] = s->m_Positions[i].y; }
        if (vel) { vel[2 * i] = s->m_Velocities[i].dx; vel[2 * i + 1] = s->m_Velocities[i].dy; }
        if (prefvel) { prefvel[2 * i] = s->m_PreferredVelocities[i].dx; prefvel[2 * i + 1] = s->m_PreferredVelocities[i].dy; }
        if (attraction) { attraction[2 * i] = s->m_AttractionPoints[i].x; attraction[2 * i + 1] = s->m_AttractionPoints[i].y; }
        if (force) { force[2 * i] = s->m_Forces[i].dx; force[2 * i + 1] = s->m_Forces[i].dy; }
        if (active) active[i] = s->m_ActiveAgents[i] ? 1 : 0;
    }
}

// ECM::GetECMCell (ECM.cpp:220-223) for arbitrary points; -1 when the reference returns nullptr.
void ecmref_query_cells(void* h, int n, const float* xy, int* out_cell) {
    RefSim* r = (RefSim*)h;
    const ECMCell* base = r->ecm->GetECMGraph().GetCells()->m_ECMCells.data();
    for (int i = 0; i < n; i++) {
        const ECMCell* c = r->ecm->GetECMCell(xy[2 * i], xy[2 * i + 1]);
        out_cell[i] = c ? (int)(c - base) : -1;
    }
}

// ECM::RetractPoint (ECM.cpp:20-96): ok flag, retracted point, edge index.
void ecmref_retract(void* h, int n, const float* xy, uint8_t* ok, float* out_xy, int* out_edge) {
    RefSim* r = (RefSim*)h;
    for (int i = 0; i < n; i++) {
        Point p;
        ECMEdge e;
        e.idx = -1;
        bool good = r->ecm->RetractPoint(Point(xy[2 * i], xy[2 * i + 1]), p, e);
        ok[i] = good ? 1 : 0;
        out_xy[2 * i] = p.x;
        out_xy[2 * i + 1] = p.y;
        out_edge[i] = e.idx;
    }
}

// Rebuilds the KD-tree on the current state (KDTree::Construct, as Simulator::Update does at
// Simulator.cpp:319) and queries every active slot in ascending order through
// Simulator::FindNNearestNeighbors with ONE shared output vector, like ORCA::m_NeighborCache
// (ORCA.h:100, ORCA.cpp:20), so stale ids carry over exactly as in the reference.
// out_ids: 5 per slot (row left untouched for inactive slots); out_counts: per slot, -1 inactive.
void ecmref_query_neighbors(void* h, int count, int* out_ids, int* out_counts) {
    RefSim* r = (RefSim*)h;
    Simulator* s = r->sim;
    s->m_KDTree->Construct(s);
    std::fill(r->nn_cache.begin(), r->nn_cache.end(), 0);  // a fresh ORCA object starts zero-filled (ORCA.h:87)
    for (int i = 0; i < count; i++) {
        out_counts[i] = -1;
        if (i > s->m_LastEntityIdx || !s->m_ActiveAgents[i]) continue;
        int nn = 0;
        s->FindNNearestNeighbors(i, 5, r->nn_cache, nn);
        out_counts[i] = nn;
        for (int k = 0; k < 5; k++) out_ids[5 * i + k] = r->nn_cache[k];
    }
}

// Simulator::FindNearestObstacles (Simulator.cpp:259-292) with ORCA's range (ORCA.cpp:27);
// obstacle vertices are reported as flat indices in (obstacle, vertex) order.
int ecmref_query_obstacles(void* h, int slot, int* out_ids, int cap) {
    RefSim* r = (RefSim*)h;
    Simulator* s = r->sim;
    std::unordered_map<const ObstacleVertex*, int> id;
    int n = 0;
    for (const Obstacle& o : r->env.GetObstacles())
        for (const ObstacleVertex* v : o.verts) id[v] = n++;
    float range = 10.0f * s->m_PreferredSpeed[slot].speed + s->m_Clearances[slot].clearance;
    std::vector<const ObstacleVertex*> out;
    s->FindNearestObstacles(slot, range * range, out);
    for (int k = 0; k < (int)out.size() && k < cap; k++) out_ids[k] = id[out[k]];
    return (int)out.size();
}

// ORCA::GetVelocity (ORCA.cpp:14-57) for one slot on the current state (KD-tree must be current:
// call ecmref_query_neighbors first).  Uses the simulator's own ORCA object.
void ecmref_orca_velocity(void* h, int slot, float* out_v) {
    Simulator* s = ((RefSim*)h)->sim;
    Vec2 v;
    s->m_ORCA->GetVelocity(s, slot, s->m_SimStepTime, s->m_PreferredSpeed[slot].speed, v);
    out_v[0] = v.x;
    out_v[1] = v.y;
}

}  // extern "C"
