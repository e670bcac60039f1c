// ECM construction for general polygonal environments, without Boost.
//
// The reference builds its ECM from Boost.Polygon's segment Voronoi diagram of the "obstacle union" - the
// walkable-area boundary plus every obstacle edge (ECMGenerator.cpp:235-256, Environment.cpp:188-229) - and keeps the
// primary, finite edges whose end points lie in free space (ECMGenerator.cpp:60-232).  Boost is neither vendored by
// the reference nor installed here, so the diagram is computed directly:
//
//   sites     the open segments and their end points (shared end points once);
//   vertices  every point equidistant from three sites with no site closer.  A segment site contributes the linear
//             equation  s (n.x - c) = r  (s = which side), a point site |x - p|^2 = r^2; differences of point equations
//             are linear too, so every triple reduces to a line in (x, y, r) cut with at most one quadric - one solver
//             for PPP / PPS / PSS / SSS.  Candidates are kept when each segment site is met in its interior (closed),
//             no other site is closer, and the point lies in the walkable area outside every obstacle's interior;
//   edges     for every pair of sites (a segment with its own end point excepted: Boost's "secondary" edges) the
//             vertices that hold both, ordered along their bisector - a line, or the parabola of a point and a
//             segment - and joined where the bisector point between two consecutive ones is itself nearest to the
//             pair.  Like the reference, a parabolic arc is stored by its end points only (a chord).
//
// O(n^3) triples with an O(n) emptiness test each: meant for the reference's own environments (Environment.cpp:27-185:
// a dozen segments) and scenes of up to a few hundred sites, as host-side input preparation - not for city maps, which
// the closed-form lattice generator (lattice_world.h) covers.  Output follows the conventions of flat_world.h
// (SURVEY.md Appendix A): closest points per half-edge, vertex clearance, outgoing half-edge rings.
// ECM construction parity is UNPINNED (no Boost to compare with); tests check the defining properties instead and run
// the unmodified reference's planner and simulator on the result (tests/test_polygon_world.py, golden debug1_small).
#pragma once
#include <string>
#include <vector>

#include "flat_world.h"

namespace ecmb200 {

// bbox: xmin ymin xmax ymax of the walkable area (its four edges are sites, like Environment::AddWalkableArea).
// poly_first[k] .. poly_first[k+1]: the vertices of obstacle k in poly_xy, counter-clockwise (Environment.cpp:198),
// strictly inside the area and disjoint from each other.  Returns false with a message on degenerate input.
bool BuildPolygonWorld(const float bbox[4], int n_polys, const int* poly_first, const float* poly_xy, FlatWorld& out, std::string* error = nullptr);

}  // namespace ecmb200
