"""GPU (-m gpu): the split tick (ECMGPU_SPLIT=1: k_knn_rows + k_orca_rows instead of k_orca) gives the default tick's
state bit for bit - single device and in-process strips.  tests/test_hostdev_kernels.py pins it on the CPU."""
import os

import numpy as np
import pytest

from ecmgenerator_b200 import gpu
from ecmgenerator_b200 import multigpu as M
from tests.util import GOLDEN, Golden

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300, method="thread")]


def _halo_for(g, n_strips):
    """Comfortably more than any agent's 5th-neighbour distance at the start, but no wider than the narrowest interior
    strip (ecmgpu_comm_set_strips refuses that): the recipe of tests/test_hostdev_kernels.py for the small golden scenes."""
    p = g.crowd.pos.astype(np.float64)
    d2 = ((p[:, None, :] - p[None, :, :]) ** 2).sum(-1)
    r5 = float(np.sqrt(np.sort(d2, axis=1)[:, 5]).max())
    widths = np.diff(M.strip_bounds(g.crowd.pos[:, 0], n_strips))[1:-1]
    return float(min(2.0 * r5 + 2.0, widths.min())) if len(widths) else 2.0 * r5 + 2.0


def _run(g, split, ticks):
    old = os.environ.pop("ECMGPU_SPLIT", None)
    if split:
        os.environ["ECMGPU_SPLIT"] = "1"  # read by ecmgpu_create
    try:
        sim = gpu.GpuSim(g.world, g.n + 8, g.step)
    finally:
        os.environ.pop("ECMGPU_SPLIT", None)
        if old is not None:
            os.environ["ECMGPU_SPLIT"] = old
    sim.bulk_load(g.crowd.pos, g.crowd.radius, g.crowd.speed, g.path_off, g.path_xy)
    sim.update(ticks)
    st = sim.state(g.n)
    st["nbr"] = sim.read(gpu.NEIGHBORS, 0, g.n)
    st["status"] = sim.read(gpu.STATUS, 0, g.n)
    stats = sim.stats()
    sim.close()
    return st, stats


@pytest.mark.parametrize("name", GOLDEN)
def test_split_tick_equals_default_tick_bitwise(name):
    g = Golden(name)
    ticks = min(g.ticks("exact-knn"), 120)
    a, sa = _run(g, False, ticks)
    b, sb = _run(g, True, ticks)
    for k in a:
        assert np.array_equal(a[k].view(np.uint8), b[k].view(np.uint8)), f"{name}: {k} differs"
    assert sa["lp3d_runs"] == sb["lp3d_runs"] and sa["knn_fallbacks"] == sb["knn_fallbacks"]
    assert sb["kernel_launches"] > sa["kernel_launches"]  # one more kernel per tick: the switch was on


def test_split_tick_with_strips_equals_single_device():
    g = Golden("jam_small")
    single, _ = _run(g, False, 60)
    os.environ["ECMGPU_SPLIT"] = "1"
    try:
        strips = M.LocalStrips(g.world, g.crowd, g.path_off, g.path_xy, 3, step=g.step, halo=_halo_for(g, 3))
    finally:
        os.environ.pop("ECMGPU_SPLIT", None)
    strips.update(60)
    strips.sync()
    pos, owners = strips.gather(gpu.POS)
    vel, _ = strips.gather(gpu.VEL)
    live = single["active"] > 0
    assert np.array_equal(owners > 0, live)
    pos, vel = pos[live], vel[live]
    single = {k: v[live] for k, v in single.items()}
    assert np.array_equal(pos.view(np.uint32), single["pos"].view(np.uint32))
    assert np.array_equal(vel.view(np.uint32), single["vel"].view(np.uint32))
    assert sum(s["halo_misses"] for s in strips.stats()) == 0


@pytest.mark.parametrize("name,rebalance", [("c2_small", False), ("jam_small", True)])
def test_compact_walk_strips_equal_single_device(name, rebalance):
    """ECMGPU_COMPACT=1: pack / cell count / scatter walk the owned share (device/strips.cuh WalkView)."""
    g = Golden(name)
    ticks = min(g.ticks("exact-knn"), 120)
    single, _ = _run(g, False, ticks)
    os.environ["ECMGPU_COMPACT"] = "1"
    try:
        strips = M.LocalStrips(g.world, g.crowd, g.path_off, g.path_xy, 3, step=g.step, halo=_halo_for(g, 3))
    finally:
        os.environ.pop("ECMGPU_COMPACT", None)
    strips.update(ticks // 2)
    if rebalance:  # new borders: lists and per-rank grids are rebuilt (c2_small's middle strip is as narrow as its halo already)
        strips.rebalance()
    strips.update(ticks - ticks // 2)
    strips.sync()
    pos, owners = strips.gather(gpu.POS)
    vel, _ = strips.gather(gpu.VEL)
    live = single["active"] > 0
    assert np.array_equal(owners > 0, live)
    assert np.array_equal(pos[live].view(np.uint32), single["pos"][live].view(np.uint32))
    assert np.array_equal(vel[live].view(np.uint32), single["vel"][live].view(np.uint32))
    assert sum(s["halo_misses"] for s in strips.stats()) == 0


def test_compact_walk_in_the_graph_tick_with_spawns_and_destroys():
    """One rank that owns the whole world, driven through ecmgpu_update (the captured-graph tick): with ECMGPU_COMPACT=1
    the list-walking kernels sit inside the graph while the list itself is rebuilt outside it whenever the host changes
    who exists (a destroy leaves a stale entry, a spawn forces a rebuild).  Must equal the plain simulator doing the same."""
    g = Golden("c2_small")
    n0 = g.n - 20  # the last 20 agents are spawned later

    def run(compact):
        os.environ.pop("ECMGPU_COMPACT", None)
        if compact:
            os.environ["ECMGPU_COMPACT"] = "1"
        try:
            sim = gpu.GpuSim(g.world, g.n + 8, g.step)
        finally:
            os.environ.pop("ECMGPU_COMPACT", None)
        off = g.path_off
        sim.bulk_load(g.crowd.pos[:n0], g.crowd.radius[:n0], g.crowd.speed[:n0], off[: n0 + 1], g.path_xy[: off[n0]])
        if compact:
            x0, _, x1, _ = (float(v) for v in g.world.bbox)
            sim.comm_set_strips(np.float32([x0 - 10.0, x1 + 10.0]), 5.0)  # rank 0 of 1: no neighbours, no transport
        sim.update(15)
        for s in (3, 77, 140):
            sim.destroy_agent(s)
        sim.update(15)
        for i in range(n0, g.n):  # late arrivals, one slot at a time like Simulator::SpawnAgent
            sim.spawn(i, g.crowd.pos[i], g.crowd.radius[i], g.crowd.speed[i], g.path_xy[off[i]:off[i + 1]])
        sim.update(30)
        st = sim.state(g.n)
        launches = sim.stats()["kernel_launches"]
        sim.close()
        return st, launches

    a, la = run(False)
    b, lb = run(True)
    assert not a["active"][[3, 77, 140]].any() and a["active"][n0:].all()
    for k in a:
        assert np.array_equal(a[k].view(np.uint8), b[k].view(np.uint8)), f"{k} differs"
    assert lb > la  # the pack kernel (and the list rebuilds) ran
