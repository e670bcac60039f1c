"""CPU: structural invariants of the analytic lattice ECM (csrc/host/lattice_world.cpp) in the
reference's conventions, checked with the C oracle's point location."""
import numpy as np
import pytest

from ecmgenerator_b200 import scenarios as S
from ecmgenerator_b200.host import lattice_world
from oracle.pyoracle import OracleSim


def test_counts_of_the_plus_world():
    # probe P6 of the survey: four corner blocks, two streets -> 17 vertices / 16 edges / 32 cells
    w = lattice_world([95, 95], [95, 95], 10.0, -100.0, -100.0)
    assert (w.n_vertices, w.n_edges, w.n_cells, w.n_obstacles) == (17, 16, 32, 4)


def test_invalid_parameters_are_rejected():
    with pytest.raises(ValueError):
        lattice_world([10], [10], 5.0)          # no street at all
    with pytest.raises(ValueError):
        lattice_world([10, 1], [10, 10], 5.0)   # block thinner than W/2
    with pytest.raises(ValueError):
        lattice_world([10, 10], [10, 10], 0.0)


@pytest.mark.parametrize("bx,by,W", [([30, 20, 25], [18, 22], 6.0), ([40] * 4, [40] * 4, 20.0), ([12], [9, 9, 9], 4.0)])
def test_half_edge_rings_and_clearances(bx, by, W):
    w = lattice_world(bx, by, W)
    nE = w.n_edges
    # every half-edge's `next` leaves the same source vertex; rings are closed
    src = np.empty(2 * nE, np.int32)
    src[0::2], src[1::2] = w.edge_v[:, 0], w.edge_v[:, 1]
    assert (src[w.he_next] == src).all()
    assert (src[w.vert_he] == np.arange(w.n_vertices)).all()
    deg = np.bincount(src, minlength=w.n_vertices)
    for v in range(w.n_vertices):
        h, seen = int(w.vert_he[v]), 0
        while True:
            seen += 1
            h = int(w.he_next[h])
            if h == w.vert_he[v]:
                break
            assert seen <= 8
        assert seen == deg[v]
    # vertex clearance = distance to its closest points
    d0 = np.linalg.norm(w.edge_cl[:, 0] - w.vert_xy[w.edge_v[:, 0]], axis=1)
    d1 = np.linalg.norm(w.edge_cl[:, 2] - w.vert_xy[w.edge_v[:, 1]], axis=1)
    assert np.allclose(d0, w.vert_clear[w.edge_v[:, 0]], atol=1e-4)
    assert np.allclose(d1, w.vert_clear[w.edge_v[:, 1]], atol=1e-4)
    # left/right really are left/right of the edge direction
    d = w.vert_xy[w.edge_v[:, 1]] - w.vert_xy[w.edge_v[:, 0]]
    for k, sign in ((0, 1), (1, -1)):
        r = w.edge_cl[:, k] - w.vert_xy[w.edge_v[:, 0]]
        cross = d[:, 0] * r[:, 1] - d[:, 1] * r[:, 0]
        assert (sign * cross > 0).all()
    # obstacles: CCW boxes, all convex
    assert w.obst_convex.all()
    assert (w.obst_next[w.obst_prev] == np.arange(w.n_obst_vertices)).all()


def test_cells_tile_the_free_space():
    w = lattice_world([30, 20, 25], [18, 22, 16], 6.0, 3.0, -7.0)
    o = OracleSim(w, 4, 1 / 60, "exact-knn")
    rng = np.random.default_rng(5)
    pts = rng.uniform(w.bbox[:2], w.bbox[2:], size=(20000, 2)).astype(np.float32)
    free = S.free_mask(w, pts, 0.01)
    blocked = ~S.free_mask(w, pts, -0.01)
    cells = o.query_cells(pts)
    assert (cells[free] >= 0).mean() > 0.999  # misses only exactly level with a vertex
    assert (cells[blocked] < 0).all()
    ok, xy, edge = o.retract(pts[free])
    assert ok.mean() > 0.999
    # retracted points lie on their edge's segment
    a, b = w.vert_xy[w.edge_v[edge[ok > 0], 0]], w.vert_xy[w.edge_v[edge[ok > 0], 1]]
    p = xy[ok > 0]
    cross = (b[:, 0] - a[:, 0]) * (p[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (p[:, 0] - a[:, 0])
    assert np.abs(cross).max() < 1e-2


def test_scenarios_are_deterministic_and_collision_free():
    w = S.world_c1()
    a, b = S.crowd_c1(w, n=800, seed=3), S.crowd_c1(w, n=800, seed=3)
    assert np.array_equal(a.pos, b.pos) and np.array_equal(a.goal, b.goal)
    d = np.linalg.norm(a.pos[:, None] - a.pos[None], axis=2) + np.eye(800) * 1e9
    assert d.min() > 0.6  # 2 * radius
    assert S.free_mask(w, a.pos, 0.3).all() and S.free_mask(w, a.goal, 0.3).all()
