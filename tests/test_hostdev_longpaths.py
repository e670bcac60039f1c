"""CPU: the polyline scan of k_attract on LONG routes (several blocks of 8 segments: block bounding boxes, backward
scan) with agents anywhere along their routes, against the C
oracle, which evaluates every segment in order like IRMPathFollower::FindAttractionPoint (IRMPathFollower.cpp:54-112).
The golden scenes' routes mostly fit one block; this is the test of the multi-block paths.  Test infrastructure only."""
import numpy as np
import pytest

from ecmgenerator_b200 import host
from ecmgenerator_b200 import scenarios as S
from oracle.pyoracle import OracleSim
from tests.test_hostdev_kernels import EmuDevice, load_emu
from tests.util import assert_bits_equal


class _Scene:
    def __init__(self, world, crowd, off, pxy, step):
        self.world, self.crowd, self.path_off, self.path_xy, self.step, self.n = world, crowd, off, pxy, step, crowd.n


def test_attraction_points_on_long_routes():
    variant = "default"
    emu = load_emu()
    w = S.world_c3()
    n = 700
    c = S.sample_crowd(w, n, 17, window=(-600, -600, 600, 600), min_goal_dist=700.0)
    off, pxy, _ = host.plan_paths(w, c.pos, c.goal, c.radius, threads=0)
    keep = np.flatnonzero(np.diff(off) >= 2)
    c = c.take(keep)
    off, pxy, _ = host.plan_paths(w, c.pos, c.goal, c.radius, threads=0)
    lens = np.diff(off)
    assert (lens >= 2).all() and np.median(lens) > 17 and (lens > 41).sum() > 20, "routes must span several blocks of 8 segments, some more than five"
    # put every agent somewhere ALONG its route (with a small offset), so the producing block is anywhere in the route
    rng = np.random.default_rng(5)
    pos = c.pos.copy()
    for i in range(c.n):
        p = pxy[off[i]:off[i + 1]]
        k = rng.integers(0, len(p) - 1)
        t = rng.random()
        pos[i] = p[k] + t * (p[k + 1] - p[k]) + rng.normal(0, 0.6, 2)
    pos = pos.astype(np.float32)
    c.pos[:] = pos
    g = _Scene(w, c, off, pxy, float(S.DT))
    d = EmuDevice(emu, g, 200.0)  # a sparse crowd: eight rings of 200 m cells reach every 5th neighbour
    ora = OracleSim(w, c.n + 8, g.step, "exact-knn")
    ora.bulk_load(pos, c.radius, c.speed, off, pxy)
    for t in range(3):
        assert emu.emu_tick(d.h) == 0
        ora.step(1)
        a, b = d.state(), ora.state(c.n)
        assert np.array_equal(a["active"], b["active"])
        for k in ("attraction", "prefvel", "pos", "vel"):
            assert_bits_equal(a[k], b[k], f"{variant}: {k} after tick {t}")
    located = d.state()["cell"] >= 0
    print(f"{variant}: {c.n} agents, routes of {int(np.median(lens))} points (max {lens.max()}), {int(located.sum())} located")
    assert located.mean() > 0.5
    d.close()
    ora.close()


def test_nan_obstacle_constraint_case_on_the_device_code():
    """tests/golden/nan_case.npz (see tests/test_oracle_golden.py): the device code must pass over the NaN obstacle
    constraint exactly like the reference's RandomizedLP does (ORCA.cpp:499-507), not project on it."""
    import os

    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nan_case.npz"))
    emu = load_emu()
    w = S.world_c3()
    m = len(z["near"])
    c = S.Crowd(z["pos"].copy(), z["pos"].copy(), z["radius"].copy(), z["speed"].copy())
    g = _Scene(w, c, z["path_off"], z["path_xy"], float(S.DT))
    d = EmuDevice(emu, g, 3.4)
    d.set_state(vel=z["vel"], attraction=z["attraction"])
    emu.emu_tick(d.h)
    a = d.state()
    assert np.isfinite(a["vel"]).all()
    for k in ("attraction", "prefvel", "pos", "vel"):
        assert_bits_equal(a[k], z["ref_" + k], k)
    d.close()
