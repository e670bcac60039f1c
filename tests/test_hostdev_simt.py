"""CPU: the kernels under a SIMT emulator (tests/hostdev/shim/simt.h, -DHD_SIMT): every CTA runs as real blocks of
co-operating threads (fibers) - 32-lane warps whose shuffles, ballots and max-reductions exchange values, CTA barriers,
CTA-wide reservations - with the launch geometry of the product.  What the one-thread-per-block emulation cannot
exercise is exercised here: the three scan kernels, the warp-synchronous loops with their real trip counts (lanes
with different bounds in lock-step), the phase barriers of k_orca, the 1024-thread pack / collect kernels, the
warp-per-agent exhaustive neighbour search of k_fallback, the carried-list kernel of the KD-tree mode.  A barrier that
some lanes never reach aborts the run.  Bit for bit against the reference's golden trajectories.  Test infrastructure."""
import ctypes as C

import numpy as np
import pytest

from ecmgenerator_b200 import multigpu as M
from ecmgenerator_b200.gpu import AGENT_REC
from tests.test_hostdev_kernels import (C_TOTAL_HALO_MISS, C_TOTAL_LP3D, EmuDevice, EmuStrips, _cell_for, _r5_max, _run_against_golden,
                                        load_emu)
from tests.util import GOLDEN, Golden, assert_bits_equal

C_TOTAL_FALLBACK = 4  # enum Counter (csrc/device/tick.cuh)


@pytest.fixture(scope="module")
def simt():
    return load_emu(["-DHD_SIMT"], "_simt")


@pytest.mark.parametrize("name", GOLDEN)
def test_simt_kernels_reproduce_reference_trajectories_bitwise(simt, name):
    g = Golden(name)
    d = EmuDevice(simt, g, _cell_for(g))
    # the jam runs in full (its LP3D queue and finisher fill up late); the two calm scenes show nothing new after a while
    _run_against_golden(d, g, lambda: simt.emu_tick(d.h), d.state, f"{name} / simt", max_ticks=None if name == "jam_small" else 24)
    if name == "jam_small":
        assert d.counters()[C_TOTAL_LP3D] > 300
    d.close()


def test_simt_exhaustive_fallback_search(simt):
    """A grid cell far too small for the crowd: eight rings do not reach the 5th neighbour, every agent goes through
    the warp-per-agent exhaustive search and its shuffle merge (knn_exhaustive) - and the trajectory stays the same."""
    g = Golden("c2_small")
    d = EmuDevice(simt, g, 0.12)
    fallbacks = 0
    mode = "exact-knn"
    ticks = 1
    for t in range(ticks):
        fallbacks += simt.emu_tick(d.h)
        st = d.state()
        assert_bits_equal(st["pos"], g.z[f"{mode}/pos"][t], f"pos after tick {t}")
        assert_bits_equal(st["vel"], g.z[f"{mode}/vel"][t], f"vel after tick {t}")
        if t == 0:
            assert_bits_equal(st["nbr"], g.z[f"{mode}/nbr0_ids"], "neighbour ids of tick 0")
    print(f"{fallbacks} exhaustive searches in {ticks} ticks of {g.n} agents")
    assert fallbacks > 0.9 * ticks * g.n and int(d.counters()[C_TOTAL_FALLBACK]) == fallbacks
    d.close()


@pytest.mark.parametrize("compact", [0, 1])
def test_simt_three_strips(simt, compact):
    """k_pack / k_pack_walk as 1024-thread CTAs with CTA-wide reservations, ghosts, migration."""
    g = Golden("jam_small")
    r5 = _r5_max(g)
    widths = np.diff(M.strip_bounds(g.crowd.pos[:, 0], 3))[1:-1]
    halo = float(min(2.0 * r5 + 2.0, widths.min()))
    s = EmuStrips(simt, g, _cell_for(g), 3, halo, narrow_grid=bool(compact))
    for dev in s.devs:
        simt.emu_set_compact(dev.h, compact)
    st = _run_against_golden(s, g, s.step, s.state, f"jam_small / simt, 3 strips, compact={compact}", max_ticks=24)
    assert st["owners"].max() == 1
    assert sum(int(d.counters()[C_TOTAL_HALO_MISS]) for d in s.devs) == 0
    seen = np.zeros(g.n, np.int32)
    for d in s.devs:  # k_collect_owned(_walk): 1024-thread CTAs
        rec = np.zeros(g.n, AGENT_REC)
        m = simt.emu_collect_owned(d.h, rec.ctypes.data_as(C.c_void_p))
        seen[rec[:m]["slot"]] += 1
    assert np.array_equal(seen, (st["active"] > 0).astype(np.int32))
    s.close()


def test_simt_kd_mode(simt):
    g = Golden("c2_small")
    d = EmuDevice(simt, g, _cell_for(g))
    simt.emu_kd_reset(d.h)

    def step():
        simt.emu_tick_kd(d.h)
        return 0

    _run_against_golden(d, g, step, d.state, "c2_small / simt kd", mode="ref-kdtree", max_ticks=32)
    d.close()


def test_simt_carried_list_kernel(simt):
    """k_kd_cache with 256 co-operating threads (the one-thread emulation once hid that only thread 0 copied)."""
    from tests.test_hostdev_kernels import _p, i32p, u8p

    simt.emu_kd_resolve.argtypes = [C.c_int, u8p, i32p, i32p, i32p, i32p, i32p]
    rng = np.random.default_rng(9)
    for n, last in ((1000, 999), (1000, 3), (700, 511), (300, 256)):
        active = np.zeros(n, np.uint8)
        active[rng.integers(0, last + 1, size=max(1, last // 3))] = 1
        active[last] = 1
        active[last + 1:] = 0
        raw = rng.integers(0, n, size=(n, 5)).astype(np.int32)
        cnt = np.full(n, 5, np.int32)
        cache = np.full(5, -5, np.int32)
        nbr, nbr_cnt = np.full((n, 5), -9, np.int32), np.full(n, -9, np.int32)
        simt.emu_kd_resolve(n, _p(active, u8p), _p(raw, i32p), _p(cnt, i32p), _p(cache, i32p), _p(nbr, i32p), _p(nbr_cnt, i32p))
        assert np.array_equal(cache, raw[last]), (n, last)


def test_simt_lp3d_of_a_packed_crowd(simt):
    """A crowd packed into a few metres, everybody heading elsewhere: hundreds of infeasible 2-D programs per tick, with
    obstacle constraints among them.  k_orca parks those agents, k_fallback runs RandomizedLP3D for a warp's 32 agents in
    step (randomized_lp3d_warp) - bit for bit the C oracle, which runs the reference's loops in place.  (The golden scenes
    reach the queue with a few agents per tick; this fills it.)"""
    from ecmgenerator_b200 import host
    from oracle.pyoracle import OracleSim

    class _Scene:
        def __init__(self, world, crowd, off, pxy, step):
            self.world, self.crowd, self.path_off, self.path_xy, self.step, self.n = world, crowd, off, pxy, step, crowd.n

    g0 = Golden("concave_small")  # recessed, turned blocks: obstacle constraints with concave ends in the programs
    w = g0.world
    rng = np.random.default_rng(11)
    n = 640
    centre = g0.crowd.pos.mean(axis=0)
    c = g0.crowd.take(rng.integers(0, g0.n, n))
    side = int(np.ceil(np.sqrt(n)))
    lattice = np.stack(np.meshgrid(np.arange(side), np.arange(side)), -1).reshape(-1, 2)[:n].astype(np.float32)
    c.pos[:] = (centre + (lattice - side / 2) * 0.62 + rng.normal(0, 0.02, (n, 2))).astype(np.float32)
    off, pxy, _ = host.plan_paths(w, c.pos, c.goal, c.radius, threads=0)
    c = c.take(np.flatnonzero(np.diff(off) >= 2))
    off, pxy, _ = host.plan_paths(w, c.pos, c.goal, c.radius, threads=0)
    assert c.n > 200 and (np.diff(off) >= 2).all()
    g = _Scene(w, c, off, pxy, g0.step)
    d = EmuDevice(simt, g, 2.0)
    ora = OracleSim(w, c.n + 8, g.step, "exact-knn")
    ora.bulk_load(c.pos, c.radius, c.speed, off, pxy)
    for t in range(12):
        simt.emu_tick(d.h)
        ora.step(1)
        a, b = d.state(), ora.state(c.n)
        for k in ("pos", "vel"):
            assert_bits_equal(a[k], b[k], f"{k} after tick {t}")
    runs = int(d.counters()[C_TOTAL_LP3D])
    print(f"{c.n} agents, {runs} LP3D runs in 12 ticks; oracle counters {ora.counters()}")
    assert runs > 400
    d.close()
    ora.close()
