"""GPU (-m gpu): the compact walk of strips (default; ECMGPU_COMPACT=0 restores the all-slots walk) gives the
single-device state bit for bit - in-process strips and the captured-graph tick.  tests/test_hostdev_kernels.py pins
it on the CPU."""
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


def _run(g, ticks):
    sim = gpu.GpuSim(g.world, g.n + 8, g.step)
    sim.bulk_load(g.crowd.pos, g.crowd.radius, g.crowd.speed, g.path_off, g.path_xy)
    sim.update(ticks)
    st = sim.state(g.n)
    st["nbr"] = sim.read(gpu.NEIGHBORS, 0, g.n)
    st["status"] = sim.read(gpu.STATUS, 0, g.n)
    stats = sim.stats()
    sim.close()
    return st, stats


@pytest.mark.parametrize("name,rebalance,compact", [("c2_small", False, 1), ("jam_small", True, 1), ("jam_small", True, 0)])
def test_compact_walk_strips_equal_single_device(name, rebalance, compact):
    """pack / cell count / scatter walk the owned share (device/strips.cuh WalkView); ECMGPU_COMPACT=0: every slot."""
    g = Golden(name)
    ticks = min(g.ticks("exact-knn"), 120)
    single, _ = _run(g, ticks)
    os.environ["ECMGPU_COMPACT"] = str(compact)
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
    """One rank that owns the whole world, driven through ecmgpu_update (the captured-graph tick): with the compact walk
    the list-walking kernels sit inside the graph while the list itself is rebuilt outside it whenever the host changes
    who exists (a destroy leaves a stale entry, a spawn forces a rebuild).  Must equal the plain simulator doing the same."""
    g = Golden("c2_small")
    n0 = g.n - 20  # the last 20 agents are spawned later

    def run(compact):
        sim = gpu.GpuSim(g.world, g.n + 8, g.step)  # the walk only exists with strips: `compact` = one strip over everything
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
