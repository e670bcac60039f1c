// Value types of the drop-in Simulator (csrc/dropin/Simulator.h): the per-slot component structs a host reads
// through the getters, and the spawn / goal / obstacle areas.  Names, fields and layouts are the reference's
// (ECMAgentSimulator/Simulator.h:29-57 components, ECMAgentSimulator/Area.h:11-105 areas, ECMDataTypes.h Point / Vec2)
// because host code written against the reference indexes these structs directly.
#pragma once
#include <vector>

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

}  // namespace Simulation
}  // namespace ECM
