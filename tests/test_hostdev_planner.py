"""CPU: the device path planner (csrc/device/planner.cuh) compiled for the host and run through
the kernel's worker loop (tests/hostdev/kernels_emul.cpp).  Checked bit for bit against
  * the REFERENCE planner's own polylines stored in the golden files (crowd/path_*: ECMPathPlanner::FindPath as
    Simulator::SpawnAgent called it for every agent of the three scenes, ~1000 queries), and
  * csrc/host/planner.cpp on thousands of random queries in larger worlds, failures included (no cell, impassable
    clearance, start and goal on one edge).
Test infrastructure only."""
import ctypes as C

import numpy as np
import pytest

from ecmgenerator_b200 import host
from ecmgenerator_b200 import scenarios as S
from tests.test_hostdev_kernels import EmuDevice, _p, emu, f32p, i32p, u8p  # noqa: F401
from tests.util import GOLDEN, Golden, assert_bits_equal


class _Scene:
    def __init__(self, world, step=1 / 60):
        self.world, self.step, self.n = world, step, 4
        self.crowd = S.Crowd(np.zeros((4, 2), np.float32), np.zeros((4, 2), np.float32), np.full(4, 0.3, np.float32), np.full(4, 1.4, np.float32))
        self.path_off = np.arange(0, 10, 2, dtype=np.int32)
        self.path_xy = np.zeros((8, 2), np.float32)
        self.crowd.pos[:] = self.path_xy[::2]


def _emu_plan(emu, world, start, goal, clearance, workers=7, cap_path=2048, cap_portals=8192, cap_out=1024, pool_cap=None, dev=None, cap_push=0):
    """Capacities default to the second (full) pass of ecmgpu.cu's plan_alloc; cap_push = 0: room for every push (2E + 4)."""
    own = dev is None
    if own:
        dev = EmuDevice(emu, _Scene(world), 4.0)
    w = world
    emu.emu_set_topology.argtypes = [C.c_void_p, f32p, i32p, i32p]
    emu.emu_plan_paths.argtypes = [C.c_void_p, C.c_int, C.c_int, f32p, f32p, f32p, i32p, i32p, u8p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    keep = [np.ascontiguousarray(w.vert_clear, np.float32), np.ascontiguousarray(w.vert_he, np.int32), np.ascontiguousarray(w.he_next, np.int32)]
    emu.emu_set_topology(dev.h, _p(keep[0], f32p), _p(keep[1], i32p), _p(keep[2], i32p))
    n = len(start)
    a = [np.ascontiguousarray(x, np.float32) for x in (start, goal, clearance)]
    off, ln, st = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.uint8)
    pool_cap = pool_cap or 256 * n + 64
    pool = np.zeros((pool_cap, 2), np.float32)
    used = emu.emu_plan_paths(dev.h, workers, n, _p(a[0], f32p), _p(a[1], f32p), _p(a[2], f32p), _p(off, i32p), _p(ln, i32p), _p(st, u8p),
                              _p(pool, f32p), pool_cap, cap_path, cap_portals, cap_out, cap_push)
    assert used >= 0, "a query left the A* arrays dirty"
    if own:
        dev.close()
    return off, ln, st, pool, used


def _compare(off, ln, st, pool, ref_off, ref_xy, label):
    ref_len = np.diff(ref_off)
    assert np.array_equal(ln, ref_len), f"{label}: {int((ln != ref_len).sum())} path lengths differ, first {np.flatnonzero(ln != ref_len)[:5]}"
    assert np.array_equal(st == 0, ref_len > 0)
    idx = np.repeat(off.astype(np.int64), ln) + (np.arange(int(ln.sum())) - np.repeat(np.cumsum(ln) - ln, ln))
    assert_bits_equal(pool[idx], ref_xy, f"{label}: polylines")


@pytest.mark.parametrize("name", GOLDEN)
def test_device_planner_reproduces_the_reference_polylines(emu, name):
    g = Golden(name)
    off, ln, st, pool, used = _emu_plan(emu, g.world, g.crowd.pos, g.crowd.goal, g.crowd.radius)
    assert used == int(g.path_off[-1])
    _compare(off, ln, st, pool, g.path_off, g.path_xy, name)
    # the packed offsets tile the pool: no two paths overlap
    order = np.argsort(off[ln > 0])
    o, l = off[ln > 0][order], ln[ln > 0][order]
    assert (o[1:] == o[:-1] + l[:-1]).all() and o[0] == 0


@pytest.mark.parametrize("config,n", [("c2_50k", 1500), ("c3_1m", 1200), ("c5_250k", 600)])
def test_device_planner_equals_the_host_planner_on_random_queries(emu, config, n):
    world_fn, crowd_fn = S.CONFIGS[config]
    w = world_fn()
    c = crowd_fn(w, n=n)
    rng = np.random.default_rng(8)
    start, goal, cl = c.pos.copy(), c.goal.copy(), c.radius.copy()
    x0, y0, x1, y1 = (float(v) for v in w.bbox)
    k = n // 10
    start[:k] = rng.uniform([x0, y0], [x1, y1], size=(k, 2))          # many of these lie inside blocks: no cell
    goal[k:2 * k] = start[k:2 * k] + rng.normal(0, 1.0, size=(k, 2))  # start and goal on one edge, or a neighbouring one
    cl[2 * k:3 * k] = rng.uniform(2.0, 0.6 * float(w.street_width), size=k)  # wide agents: parts of the graph impassable
    cl[3 * k:3 * k + 20] = 50.0                                       # nothing is passable
    ref_off, ref_xy, n_ok = host.plan_paths(w, start, goal, cl, threads=0)
    off, ln, st, pool, used = _emu_plan(emu, w, start, goal, cl, workers=13)
    print(f"{config}: {n_ok} of {n} queries have a path, {int(ref_off[-1])} points")
    assert 0.5 * n < n_ok < n
    _compare(off, ln, st, pool, ref_off, ref_xy, config)


def test_device_planner_reports_capacity_overflow(emu):
    g = Golden("c2_small")
    ref_len = np.diff(g.path_off)
    off, ln, st, pool, _ = _emu_plan(emu, g.world, g.crowd.pos, g.crowd.goal, g.crowd.radius, cap_out=4)
    long = ref_len > 4
    assert long.any() and (st[long] == 2).all() and (ln[long] == 0).all()
    assert (st[~long] == 0).all() and np.array_equal(ln[~long], ref_len[~long])
    # a pool that is too small: the paths that did not fit are flagged, the others are intact
    off, ln, st, pool, used = _emu_plan(emu, g.world, g.crowd.pos, g.crowd.goal, g.crowd.radius, pool_cap=300)
    assert used == int(g.path_off[-1]) and (st == 2).any() and (st == 0).any()
    fit = st == 0
    assert ((off + ln)[fit] <= 300).all() and np.array_equal(ln[fit], ref_len[fit])
    # the first pass of ecmgpu_plan_paths runs with a push capacity sized for the usual query: a query that fills it is
    # reported (and planned again with the full capacity by the caller), the others are untouched by it, and the A*
    # records are left idle either way (checked inside _emu_plan)
    off, ln, st, pool, _ = _emu_plan(emu, g.world, g.crowd.pos, g.crowd.goal, g.crowd.radius, cap_push=24)
    assert (st == 2).any() and (st == 0).any()
    assert np.array_equal(ln[st == 0], ref_len[st == 0]) and (ln[st == 2] == 0).all()
    _compare(off[st == 0], ln[st == 0], st[st == 0], pool, np.concatenate([[0], np.cumsum(ref_len[st == 0])]).astype(np.int32),
             np.concatenate([g.path_xy[g.path_off[i]:g.path_off[i + 1]] for i in np.flatnonzero(st == 0)]), "small push capacity")
