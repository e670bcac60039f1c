"""GPU (-m gpu): ecmgpu_plan_paths - the batched device planner against the REFERENCE planner's
polylines in the golden files and against csrc/host/planner.cpp on a large random batch, bit for bit.
tests/test_hostdev_planner.py pins the same device code on the CPU."""
import time

import numpy as np
import pytest

from ecmgenerator_b200 import gpu, host
from ecmgenerator_b200 import scenarios as S
from tests.util import GOLDEN, Golden, assert_bits_equal

# a kernel that never returns must not hang the box: the watchdog thread ends the run instead
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600, method="thread")]


@pytest.mark.parametrize("name", GOLDEN)
def test_device_planner_reproduces_the_reference_polylines(name):
    g = Golden(name)
    sim = gpu.GpuSim(g.world, 8, g.step)
    assert sim.topology_error is None
    off, xy, n_ok = sim.plan_paths(g.crowd.pos, g.crowd.goal, g.crowd.radius)
    assert np.array_equal(off, g.path_off)
    assert_bits_equal(xy, g.path_xy, f"{name}: polylines")
    assert n_ok == g.n


@pytest.mark.parametrize("config,n", [("c2_50k", 20000), ("c3_1m", 60000)])
def test_device_planner_equals_the_host_planner(config, n):
    world_fn, crowd_fn = S.CONFIGS[config]
    w = world_fn()
    c = crowd_fn(w, n=n)
    rng = np.random.default_rng(8)
    start, goal, cl = c.pos.copy(), c.goal.copy(), c.radius.copy()
    x0, y0, x1, y1 = (float(v) for v in w.bbox)
    k = n // 10
    start[:k] = rng.uniform([x0, y0], [x1, y1], size=(k, 2))
    goal[k:2 * k] = start[k:2 * k] + rng.normal(0, 1.0, size=(k, 2))
    cl[2 * k:3 * k] = rng.uniform(2.0, 0.6 * float(w.street_width), size=k)
    cl[3 * k:3 * k + 20] = 50.0
    t0 = time.time()
    ref_off, ref_xy, ref_ok = host.plan_paths(w, start, goal, cl, threads=0)
    t_host = time.time() - t0
    sim = gpu.GpuSim(w, 8, float(S.DT))
    sim.plan_paths(start[:64], goal[:64], cl[:64])  # scratch allocation and bins outside the timing
    t0 = time.time()
    off, xy, n_ok = sim.plan_paths(start, goal, cl)
    t_dev = time.time() - t0
    print(f"{config}: {n} queries, {n_ok} paths; host planner {t_host:.2f} s (all cores), device {t_dev:.2f} s incl. transfers")
    assert n_ok == ref_ok and np.array_equal(off, ref_off)
    assert_bits_equal(xy, ref_xy, f"{config}: polylines")


def test_device_planner_small_pool_is_retried():
    g = Golden("c2_small")
    sim = gpu.GpuSim(g.world, 8, g.step)
    off, xy, _ = sim.plan_paths(g.crowd.pos, g.crowd.goal, g.crowd.radius, points_per_path=1)
    assert np.array_equal(off, g.path_off)
    assert_bits_equal(xy, g.path_xy, "polylines after the pool was enlarged")


def test_queries_that_fill_the_first_pass_scratch_are_planned_again(monkeypatch):
    """The first pass runs with scratch sized for the usual query (csrc/ecmgpu.cu: plan_alloc); a query that fills it is
    planned again with the full capacities.  With room for 24 pushes only, most routes of the golden scene take the
    second pass - and the polylines still equal the reference's bit for bit."""
    g = Golden("c2_small")
    monkeypatch.setenv("ECMGPU_PLAN_PUSH", "24")
    sim = gpu.GpuSim(g.world, 8, g.step)
    off, xy, n_ok = sim.plan_paths(g.crowd.pos, g.crowd.goal, g.crowd.radius)
    workers, ms, second = sim.plan_info()
    print(f"{second} of {g.n} queries took the second pass; {workers} workers, {ms:.3f} ms")
    assert 0 < second <= g.n
    assert np.array_equal(off, g.path_off) and n_ok == g.n
    assert_bits_equal(xy, g.path_xy, "polylines after the second pass")
    monkeypatch.delenv("ECMGPU_PLAN_PUSH")
    sim2 = gpu.GpuSim(g.world, 8, g.step)
    sim2.plan_paths(g.crowd.pos, g.crowd.goal, g.crowd.radius)
    assert sim2.plan_info()[2] == 0
