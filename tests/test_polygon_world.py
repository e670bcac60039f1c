"""CPU: the Boost-free ECM generator for polygonal scenes (csrc/host/polygon_world.cpp).

ECM construction parity is unpinned (the reference needs Boost.Polygon, which is neither vendored nor installed), so
the defining properties are checked instead - every vertex is equidistant from its three nearest obstacle features,
edges separate the right pair, rings are closed, the cells cover free space - and the UNMODIFIED reference's planner
(oracle/_ref) is run on the result: it must accept the graph and agree with the host planner bit for bit."""
import numpy as np
import pytest

from ecmgenerator_b200 import host
from ecmgenerator_b200.scenarios import SCENES, scene_polygons
from oracle import pyref


def _segments(w):
    x0, y0, x1, y1 = (float(v) for v in w.bbox)
    segs = [((x0, y0), (x1, y0)), ((x1, y0), (x1, y1)), ((x1, y1), (x0, y1)), ((x0, y1), (x0, y0))]
    for k in range(w.n_obstacles):
        p = w.obst_xy[w.obst_first[k]:w.obst_first[k + 1]].astype(np.float64)
        segs += [(tuple(p[i]), tuple(p[(i + 1) % len(p)])) for i in range(len(p))]
    return np.array(segs, np.float64)  # (m, 2, 2)


def _dist_to_segments(pts, segs):
    a, b = segs[:, 0][None], segs[:, 1][None]
    p = pts[:, None, :]
    d = b - a
    t = np.clip(((p - a) * d).sum(-1) / (d * d).sum(-1), 0.0, 1.0)
    return np.linalg.norm(p - (a + t[..., None] * d), axis=-1)  # (n, m)


def _inside_any(w, pts):
    out = np.zeros(len(pts), bool)
    for k in range(w.n_obstacles):
        poly = w.obst_xy[w.obst_first[k]:w.obst_first[k + 1]].astype(np.float64)
        x, y = pts[:, 0], pts[:, 1]
        ins = np.zeros(len(pts), bool)
        for i in range(len(poly)):
            p, q = poly[i], poly[(i + 1) % len(poly)]
            cross = ((p[1] > y) != (q[1] > y)) & (x < (q[0] - p[0]) * (y - p[1]) / (q[1] - p[1] + 1e-300) + p[0])
            ins ^= cross
        out |= ins
    return out


@pytest.mark.parametrize("name", sorted(SCENES))
def test_vertices_edges_and_rings(name):
    w = host.polygon_world(*scene_polygons(name))
    segs = _segments(w)
    v = w.vert_xy.astype(np.float64)
    d = _dist_to_segments(v, segs)
    near = np.sort(d, axis=1)
    tol = 2e-3 + 1e-5 * np.abs(w.bbox).max()
    assert np.abs(near[:, 0] - w.vert_clear).max() < tol, "clearance = distance to the nearest obstacle feature"
    # a Voronoi vertex has three nearest features (two segments meeting in a corner count as two + their common point)
    assert ((near[:, 1] - near[:, 0]) < tol).all()
    touching = w.vert_clear < tol
    assert ((near[:, 2] - near[:, 0]) < tol)[~touching].all()
    assert not _inside_any(w, v[~touching]).any()
    nE = w.n_edges
    src = np.empty(2 * nE, np.int32)
    src[0::2], src[1::2] = w.edge_v[:, 0], w.edge_v[:, 1]
    assert (src[w.he_next] == src).all() and (src[w.vert_he] == np.arange(w.n_vertices)).all()
    deg = np.bincount(src, minlength=w.n_vertices)
    for vv in range(w.n_vertices):
        h, seen = int(w.vert_he[vv]), 0
        while True:
            seen += 1
            h = int(w.he_next[h])
            if h == int(w.vert_he[vv]) or seen > 64:
                break
        assert seen == deg[vv]
    # every edge's closest points are the nearest points of ITS two sites at both ends: both at the vertex clearance
    for e in range(nE):
        for end, vi in ((0, w.edge_v[e, 0]), (1, w.edge_v[e, 1])):
            for side in (0, 1):
                c = w.edge_cl[e, 2 * end + side].astype(np.float64)
                assert abs(np.linalg.norm(c - v[vi]) - w.vert_clear[vi]) < tol
    # connected: one medial axis for one free space
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components

    adj = coo_matrix((np.ones(nE), (w.edge_v[:, 0], w.edge_v[:, 1])), shape=(w.n_vertices, w.n_vertices))
    assert connected_components(adj, directed=False)[0] == 1


@pytest.mark.parametrize("name", sorted(SCENES))
def test_cells_cover_free_space_and_paths_exist(name):
    w = host.polygon_world(*scene_polygons(name))
    rng = np.random.default_rng(3)
    bb = w.bbox.astype(np.float64)
    pts = rng.uniform(bb[:2] + 1.0, bb[2:] - 1.0, size=(6000, 2))
    free = ~_inside_any(w, pts) & (_dist_to_segments(pts, _segments(w)).min(axis=1) > 0.5)
    pts = pts[free].astype(np.float32)
    cells = host.find_cells(w, pts)
    covered = (cells >= 0).mean()
    print(f"{name}: {w.n_vertices} vertices, {w.n_edges} edges, {covered:.4f} of {len(pts)} free points located")
    assert covered > 0.97  # parabolic arcs are stored as chords, like the reference does: thin slivers stay uncovered
    ok = np.flatnonzero(cells >= 0)
    a, b = pts[ok[: 300]], pts[ok[300: 600]]
    off, xy, n_ok = host.plan_paths(w, a, b, np.full(len(a), 0.3, np.float32))
    assert n_ok > 0.95 * len(a)
    # the polylines stay in free space
    for i in range(0, len(a), 7):
        p = xy[off[i]:off[i + 1]].astype(np.float64)
        if len(p) < 2:
            continue
        samples = np.concatenate([p[j] + np.linspace(0, 1, 20)[:, None] * (p[j + 1] - p[j]) for j in range(len(p) - 1)])
        assert not _inside_any(w, samples).any()


@pytest.mark.skipif(not pyref.available("exact-knn"), reason="needs oracle/_ref (built from /root/reference)")
@pytest.mark.parametrize("name", sorted(SCENES))
def test_the_reference_planner_accepts_the_graph_and_agrees(name):
    w = host.polygon_world(*scene_polygons(name))
    rng = np.random.default_rng(5)
    bb = w.bbox.astype(np.float64)
    pts = rng.uniform(bb[:2] + 2.0, bb[2:] - 2.0, size=(1500, 2))
    free = ~_inside_any(w, pts) & (_dist_to_segments(pts, _segments(w)).min(axis=1) > 1.0)
    pts = pts[free].astype(np.float32)[:240]
    a, b = pts[:120], pts[120:240]
    off, xy, _ = host.plan_paths(w, a, b, np.full(len(a), 0.3, np.float32))
    r = pyref.RefSim(w, 8, 1 / 60, "exact-knn")
    same = 0
    for i in range(len(a)):
        ref = r.plan_path(a[i], b[i], 0.3)
        mine = xy[off[i]:off[i + 1]]
        if ref is None:
            assert len(mine) == 0
        else:
            assert np.array_equal(ref.view(np.uint32), mine.view(np.uint32)), i
            same += 1
    r.close()
    assert same > 100


def test_invalid_scenes_are_refused():
    with pytest.raises(ValueError):
        host.polygon_world((0, 0, 10, 10), [[(1, 1), (1, 3), (3, 3), (3, 1)]])  # clockwise
    with pytest.raises(ValueError):
        host.polygon_world((0, 0, 10, 10), [[(1, 1), (13, 1), (13, 3), (1, 3)]])  # leaves the area
    with pytest.raises(ValueError):
        host.polygon_world((0, 0, 0, 10), [])
