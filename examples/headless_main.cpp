// Headless host program using the drop-in ECM::Simulation::Simulator the way the reference's
// main() + Application::Run do (/root/reference/ECMApplication/main.cpp:50-54,
// Application.cpp:44-63, 137-144): build the environment's ECM, create planner and simulator, add a
// spawn and a goal area, then call Update(dt) once per frame and read the component arrays back.
//
//   ecm_headless [agents_per_second] [ticks] [device]
//
// C++17, links libecmsim.so (drop-in class + host planner) which forwards to libecmgpu.so (C ABI).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <vector>

#include "../ecmgenerator_b200/csrc/dropin/Simulator.h"
#include "../ecmgenerator_b200/csrc/host/lattice_world.h"

int main(int argc, char** argv) {
    const float rate = argc > 1 ? (float)atof(argv[1]) : 200.0f;
    const int ticks = argc > 2 ? atoi(argv[2]) : 600;
    const int device = argc > 3 ? atoi(argv[3]) : 0;

    // environment: 4 x 4 blocks of 60 x 60 with 20-wide streets, centred on the origin
    std::vector<float> blocks(4, 60.0f);
    ecmb200::FlatWorld world;
    if (!ecmb200::BuildLatticeWorld(4, blocks.data(), 4, blocks.data(), 20.0f, -150.0f, -150.0f, world)) {
        fprintf(stderr, "world construction failed\n");
        return 1;
    }
    ecmb200::PathPlanner planner(&world);
    ECM::Simulation::Simulator sim(&world, &planner, 20000, 1.0f / 60.0f, device);
    try {
        sim.Initialize();
    } catch (const std::exception& e) {
        fprintf(stderr, "Initialize failed: %s\n", e.what());
        return 2;
    }
    ECM::Simulation::SpawnConfiguration cfg;
    cfg.clearanceMin = 0.3f;
    cfg.preferredSpeedMin = 1.4f;
    // two opposing flows along the street y = -40 .. -20 ... spawn areas sit inside streets
    const int s0 = sim.AddSpawnArea(ECM::Point(-80.0f, -110.0f), ECM::Vec2(8.0f, 30.0f), cfg);
    const int g0 = sim.AddGoalArea(ECM::Point(80.0f, 110.0f), ECM::Vec2(8.0f, 30.0f));
    const int s1 = sim.AddSpawnArea(ECM::Point(80.0f, -110.0f), ECM::Vec2(8.0f, 30.0f), cfg);
    const int g1 = sim.AddGoalArea(ECM::Point(-80.0f, 110.0f), ECM::Vec2(8.0f, 30.0f));
    sim.ConnectSpawnGoalAreas(s0, g0, rate);
    sim.ConnectSpawnGoalAreas(s1, g1, rate);
    srand(1);

    const auto t0 = std::chrono::steady_clock::now();
    long long updates = 0;
    try {
        for (int t = 0; t < ticks; t++) {
            sim.Update(1.0f / 60.0f);
            updates += sim.GetNumAgents();
        }
    } catch (const std::exception& e) {
        fprintf(stderr, "Update failed: %s\n", e.what());
        return 3;
    }
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    // what the renderer would read (ECMRenderer.cpp:836-884)
    const auto* pos = sim.GetPositionData();
    const auto* vel = sim.GetVelocityData();
    const bool* act = sim.GetActiveFlags();
    double sx = 0, sy = 0, sp = 0;
    int n = 0;
    for (int i = 0; i <= sim.GetLastIndex(); i++) {
        if (!act[i]) continue;
        sx += pos[i].x; sy += pos[i].y;
        sp += std::sqrt(vel[i].dx * vel[i].dx + vel[i].dy * vel[i].dy);
        n++;
    }
    printf("ticks %d agents %d (last index %d) mean pos (%.2f, %.2f) mean speed %.3f  %.3f ms/tick %.3g agent-updates/s\n", ticks,
           sim.GetNumAgents(), sim.GetLastIndex(), n ? sx / n : 0.0, n ? sy / n : 0.0, n ? sp / n : 0.0, 1e3 * sec / ticks,
           updates / sec);
    return n == sim.GetNumAgents() ? 0 : 4;
}
