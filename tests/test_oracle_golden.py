"""CPU: the plain-C oracle (oracle/ecm_oracle.c) against golden vectors produced by the unmodified
reference (tests/golden/make_golden.py).  Bit-exact in both neighbour modes."""
import os

import numpy as np
import pytest

from oracle.pyoracle import OracleSim, _load
from tests.util import GOLDEN, Golden, apply_events, assert_bits_equal


@pytest.mark.parametrize("name", GOLDEN)
@pytest.mark.parametrize("mode", ["ref-kdtree", "exact-knn"])
def test_oracle_reproduces_reference_trajectories(name, mode):
    g = Golden(name)
    o = OracleSim(g.world, g.n + 8, g.step, mode)
    slots = o.bulk_load(g.crowd.pos, g.crowd.radius, g.crowd.speed, g.path_off, g.path_xy)
    assert (slots == np.arange(g.n)).all()
    ids, cnt = o.query_neighbors(g.n)
    assert_bits_equal(ids, g.z[f"{mode}/nbr0_ids"], "neighbour ids at t=0")
    assert_bits_equal(cnt, g.z[f"{mode}/nbr0_cnt"], "neighbour counts at t=0")
    full_at = set(int(t) for t in g.z[f"{mode}/full_at"])
    T = g.ticks(mode)
    for t in range(T):
        o.step(1)
        st = o.state(g.n)
        assert_bits_equal(st["pos"], g.z[f"{mode}/pos"][t], f"pos after tick {t}")
        assert_bits_equal(st["vel"], g.z[f"{mode}/vel"][t], f"vel after tick {t}")
        assert np.array_equal(st["active"], g.z[f"{mode}/active"][t]), f"active after tick {t}"
        if t in full_at:
            for k in ("prefvel", "attraction", "force"):
                assert_bits_equal(st[k], g.z[f"{mode}/full{t}_{k}"], f"{k} after tick {t}")
        apply_events(o, g.events_at(mode, t))
    ids, cnt = o.query_neighbors(g.n)
    assert_bits_equal(ids, g.z[f"{mode}/nbr1_ids"], "neighbour ids at the end")
    assert_bits_equal(cnt, g.z[f"{mode}/nbr1_cnt"], "neighbour counts at the end")


@pytest.mark.parametrize("name", GOLDEN)
def test_oracle_point_location_and_retraction(name):
    g = Golden(name)
    o = OracleSim(g.world, 8, g.step, "exact-knn")
    pts = g.z["probe/xy"]
    assert np.array_equal(o.query_cells(pts), g.z["probe/cell"])
    ok, xy, edge = o.retract(pts)
    assert np.array_equal(ok, g.z["probe/retract_ok"])
    good = ok > 0
    assert np.array_equal(edge[good], g.z["probe/retract_edge"][good])
    assert_bits_equal(xy[good], g.z["probe/retract_xy"][good], "retracted points")
    # the probes must exercise hits, misses and retraction failures
    assert (g.z["probe/cell"] >= 0).sum() > 100 and (g.z["probe/cell"] < 0).sum() > 100


def test_jam_golden_exercises_lp3d_and_collisions():
    g = Golden("jam_small")
    o = OracleSim(g.world, g.n + 8, g.step, "exact-knn")
    o.bulk_load(g.crowd.pos, g.crowd.radius, g.crowd.speed, g.path_off, g.path_xy)
    o.step(g.ticks("exact-knn"))
    c = o.counters()
    assert c["lp3d"] > 300, c
    assert c["max_obstacle_neighbours"] >= 4, c


def test_std_sort_restatement_handles_ties():
    """The KD-tree build depends on std::sort's order of equal keys (oracle/ecm_oracle.c, std_sort)."""
    import ctypes as C

    L = _load()
    L.eo_test_std_sort.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_float), C.c_int]
    rng = np.random.default_rng(0)
    for n in (1, 2, 16, 17, 100, 5000):
        pos = rng.integers(0, 7, size=(n, 2)).astype(np.float32)
        idx = np.arange(n, dtype=np.int32)
        L.eo_test_std_sort(idx.ctypes.data_as(C.POINTER(C.c_int)), n, pos.ctypes.data_as(C.POINTER(C.c_float)), 0)
        assert sorted(idx.tolist()) == list(range(n))
        assert (np.diff(pos[idx, 0]) >= 0).all()


@pytest.mark.parametrize("name", GOLDEN)
def test_host_locator_matches_the_reference_probes(name):
    """csrc/host/planner.cpp CellLocator (bin lists + the exact-level path) against the reference's linear scan.  In the
    turned worlds a probe level with a cell vertex is "inside" a cell far to its right under the reference's even-odd
    test (UtilityFunctions.cpp:54-86); the bin lists alone would miss that."""
    from ecmgenerator_b200 import host

    g = Golden(name)
    assert np.array_equal(host.find_cells(g.world, g.z["probe/xy"]), g.z["probe/cell"])


def test_new_goldens_hold_oblique_and_concave_geometry():
    """obstacle segments that are not axis-aligned, concave obstacle vertices, and both reaching
    ORCA::GenerateConstraints (ORCA.cpp:146-239) under the parity checks."""
    for name, want_concave in (("oblique_small", False), ("concave_small", True)):
        g = Golden(name)
        a = g.world.obst_xy
        b = a[g.world.obst_next]
        oblique = (a[:, 0] != b[:, 0]) & (a[:, 1] != b[:, 1])
        assert oblique.all(), name
        assert ((g.world.obst_convex == 0).sum() > 0) == want_concave, name
        o = OracleSim(g.world, g.n + 8, g.step, "exact-knn")
        o.bulk_load(g.crowd.pos, g.crowd.radius, g.crowd.speed, g.path_off, g.path_xy)
        o.step(g.ticks("exact-knn"))
        c = o.counters()
        assert c["oblique_segments"] > 50_000, c
        if want_concave:
            assert c["concave_segments"] > 100_000 and c["lp3d"] > 300, c


def test_nan_obstacle_constraint_is_passed_over_like_the_reference():
    """tests/golden/nan_case.npz: 486 agents around one that is exactly level with a block corner it touches (found at tick
    832 of the 1 M-agent run).  ORCA.cpp:171-183 then takes the square root of a negative number: a NaN obstacle
    constraint, which RandomizedLP passes over because `if (d <= 0) return i; else if (d > 0) {...}` (ORCA.cpp:499-507)
    does neither for NaN.  The expected state in the file was computed by the UNMODIFIED reference (oracle/_ref, both
    neighbour modes agree); a restatement that projects on the NaN constraint turns the agent's velocity into NaN."""
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nan_case.npz"))
    from ecmgenerator_b200 import scenarios as S

    w = S.world_c3()
    m = len(z["near"])
    o = OracleSim(w, m + 8, float(S.DT), "exact-knn")
    o.bulk_load(z["pos"], z["radius"], z["speed"], z["path_off"], z["path_xy"])
    for k in range(m):
        o.set_kinematics(k, z["pos"][k], z["vel"][k])
        o.set_attraction(k, z["attraction"][k])
    o.step(1)
    st = o.state(m)
    assert np.isfinite(st["vel"]).all()
    for k in ("pos", "vel", "prefvel", "attraction", "force"):
        assert_bits_equal(st[k], z["ref_" + k], k)
