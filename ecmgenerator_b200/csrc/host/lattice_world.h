// Analytic ECM construction for lattice ("city block") worlds.
//
// The reference builds its ECM from a Boost.Polygon segment Voronoi diagram
// (/root/reference/ECMGenerator/ECMGenerator.cpp:60-256).  Boost is neither vendored by the
// reference nor installed here, and ECM construction is host-side input preparation, not part of
// the per-tick hot path.  For worlds made of axis-aligned rectangular blocks separated by streets
// of one uniform width W the medial axis is known in closed form, so we emit it directly, in the
// reference's conventions (validated against the unmodified reference by
// probe P6):
//
//   * every street is a straight centre line (segment/segment bisector, clearance W/2);
//   * every crossing is a W x W square whose four nearest sites are the four block corners:
//     a centre vertex (clearance W/sqrt 2) joined to four "mouth" vertices (clearance W/2) by
//     point/point bisectors, cells degenerate to triangles;
//   * every street dead-ends at the outer wall: the centre line stops W/2 short of the wall at a
//     vertex T and two 45-degree bisectors run from T into the two wall corners (clearance 0).
//
// Blocks on the rim touch the outer wall, so the free space is exactly the union of the streets.
#pragma once
#include "flat_world.h"

namespace ecmb200 {

// bx[0..nbx) / by[0..nby): block extents along x / y.  Streets of width W lie between consecutive
// blocks: (nbx-1) vertical and (nby-1) horizontal streets.  (x0,y0) is the lower-left world corner.
// Returns false (and leaves `out` untouched) on invalid input (W <= 0, a block thinner than W/2, ...).
bool BuildLatticeWorld(int nbx, const float* bx, int nby, const float* by, float W, float x0, float y0,
                       FlatWorld& out);

}  // namespace ecmb200
