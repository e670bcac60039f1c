"""GPU (-m gpu): the faithful KD-tree neighbour mode (ecmgpu_set_neighbor_mode(ECMGPU_NEIGHBORS_KDTREE), SURVEY.md §8
row f1) through the C ABI against the UNMODIFIED reference: its golden neighbour lists and trajectories
("ref-kdtree": the reference's own KDTree.cpp, over-pruning, duplicated and stale ids included) and the C oracle in
the same mode.  Bars: neighbour lists id for id; velocities within 1e-4 m/s per step on identical input state.

A tick whose tree has a tie at a segment's median is std::sort-defined in the reference itself; the kernels count
such ticks (stats()["kd_median_ties"]) and the comparisons end there.  tests/test_hostdev_kdtree.py pins the same
device code bit for bit on the CPU.  (The file name sorts last on purpose: this mode is newer than the default path.)
"""
import numpy as np
import pytest

from ecmgenerator_b200 import gpu
from oracle.pyoracle import OracleSim
from tests.util import GOLDEN, Golden, apply_events, assert_bits_equal

# a kernel that never returns must not hang the box: the watchdog thread ends the run instead
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600, method="thread")]

VEL_TOL = 1e-4  # m/s absolute per step (north_star)
MODE = "ref-kdtree"


def _kd_sim(g):
    sim = gpu.GpuSim(g.world, g.n + 8, g.step)
    sim.bulk_load(g.crowd.pos, g.crowd.radius, g.crowd.speed, g.path_off, g.path_xy)
    sim.set_neighbor_mode(gpu.NEIGHBORS_KDTREE)
    return sim


@pytest.mark.parametrize("name", GOLDEN)
def test_kd_neighbour_lists_equal_the_unmodified_reference(name):
    g = Golden(name)
    sim = _kd_sim(g)
    ids, cnt = sim.query_neighbors(g.n)
    assert sim.stats()["kd_median_ties"] == 0
    assert_bits_equal(cnt, g.z[f"{MODE}/nbr0_cnt"], "neighbour counts")
    assert_bits_equal(ids, g.z[f"{MODE}/nbr0_ids"], "neighbour ids")
    # and they are NOT the exact 5-NN lists the default mode returns
    sim.set_neighbor_mode(gpu.NEIGHBORS_EXACT)
    ids_x, _ = sim.query_neighbors(g.n)
    assert_bits_equal(ids_x, g.z["exact-knn/nbr0_ids"], "exact lists after switching back")
    assert (np.sort(ids, 1) != np.sort(ids_x, 1)).any(1).mean() > 0.1


@pytest.mark.parametrize("name", GOLDEN)
def test_kd_lockstep_velocities_within_tolerance(name):
    """Every tick starts from the ORACLE's state (reference KD-tree mode) on both sides, then one tick each."""
    g = Golden(name)
    sim = _kd_sim(g)
    ora = OracleSim(g.world, g.n + 8, g.step, MODE)
    ora.bulk_load(g.crowd.pos, g.crowd.radius, g.crowd.speed, g.path_off, g.path_xy)
    n, worst, exact_rows, total_rows, done = g.n, 0.0, 0, 0, 0
    for t in range(g.ticks(MODE)):
        st = ora.state(n)
        sim.write(gpu.POS, st["pos"])
        sim.write(gpu.VEL, st["vel"])
        sim.write(gpu.ATTRACTION, st["attraction"])
        sim.write(gpu.ACTIVE, st["active"])
        ora.step(1)
        sim.step(1)
        if sim.stats()["kd_median_ties"] > 0:
            break
        a, b = sim.state(n), ora.state(n)
        act = st["active"] > 0
        assert np.array_equal(a["active"], b["active"]), f"active flags after tick {t}"
        alive = act & (b["active"] > 0)
        assert_bits_equal(a["prefvel"][alive], b["prefvel"][alive], f"prefvel tick {t}")
        dv = float(np.abs(a["vel"][alive] - b["vel"][alive]).max())
        worst = max(worst, dv)
        assert dv <= VEL_TOL, f"tick {t}: max |dv| = {dv}"
        exact_rows += int((a["vel"][alive].view(np.uint32) == b["vel"][alive].view(np.uint32)).all(axis=1).sum())
        total_rows += int(alive.sum())
        apply_events(ora, g.events_at(MODE, t))
        apply_events(sim, g.events_at(MODE, t))
        sim.write(gpu.REPLAN_PENDING, np.zeros(n, np.uint8))  # every failed tick is answered (Simulator.cpp:581-587), changed path or not
        done += 1
    print(f"{name}: {done} ticks in lockstep, worst |dv| {worst:.3e}; {exact_rows}/{total_rows} velocity rows bit-identical")
    assert done >= min(96, g.ticks(MODE)) and exact_rows / total_rows > 0.75


@pytest.mark.parametrize("name", GOLDEN)
def test_kd_free_running_matches_the_unmodified_reference(name):
    """No re-synchronisation: GPU trajectories in KD-tree mode vs the unmodified reference's golden trajectories."""
    g = Golden(name)
    sim = _kd_sim(g)
    n, rms, done = g.n, [], 0
    for t in range(g.ticks(MODE)):
        sim.step(1)
        if sim.stats()["kd_median_ties"] > 0:
            break
        pos = sim.read(gpu.POS, 0, n)
        act = g.z[f"{MODE}/active"][t] > 0
        assert np.array_equal(sim.read(gpu.ACTIVE, 0, n) > 0, act)
        d = pos[act] - g.z[f"{MODE}/pos"][t][act]
        rms.append(float(np.sqrt((d ** 2).sum(axis=1).mean())))
        if t == 0:  # the lists the first tick used (carried list still zero-filled)
            live = act
            assert_bits_equal(sim.read(gpu.NEIGHBORS, 0, n)[live], g.z[f"{MODE}/nbr0_ids"][live], "lists of tick 0")
        apply_events(sim, g.events_at(MODE, t))
        sim.write(gpu.REPLAN_PENDING, np.zeros(n, np.uint8))
        done += 1
    print(f"{name}: {done} ticks, trajectory RMS divergence {rms[-1]:.3e} (max {max(rms):.3e})")
    assert done >= min(96, g.ticks(MODE)) and max(rms) < 1e-2
    # the exact-kNN trajectory is a different one: this mode follows the reference's, not ours
    other = g.z["exact-knn/pos"][done - 1]
    both = (g.z[f"{MODE}/active"][done - 1] > 0) & (g.z["exact-knn/active"][done - 1] > 0)
    gap = float(np.sqrt(((g.z[f"{MODE}/pos"][done - 1][both] - other[both]) ** 2).sum(axis=1).mean()))
    print(f"{name}: RMS gap between the reference's KD-tree and exact-kNN trajectories at that tick: {gap:.3e}")


def test_kd_mode_refuses_strips():
    g = Golden("c2_small")
    from ecmgenerator_b200 import multigpu as M

    strips = M.LocalStrips(g.world, g.crowd, g.path_off, g.path_xy, 2, step=g.step, halo=10.0)
    with pytest.raises(gpu.EcmGpuError):
        strips.sims[0].set_neighbor_mode(gpu.NEIGHBORS_KDTREE)


def test_dropin_in_kd_mode_walks_like_the_unmodified_reference():
    """The C++ drop-in Simulator against the UNMODIFIED reference Simulator (its own KDTree.cpp, ORCA.cpp, ...), both
    driven through the same calls - SpawnAgent, Update, DestroyAgent, the getters - for 120 free-running ticks."""
    from ecmgenerator_b200 import dropin
    from ecmgenerator_b200 import scenarios as S
    from oracle import pyref

    if not pyref.available(MODE):
        pytest.skip("oracle/_ref not built")
    w = S.world_c1()
    c = S.crowd_c1(w, n=400, seed=61)
    ref = pyref.RefSim(w, 512, 1 / 60, MODE)
    sim = dropin.Simulator(w, 512, 1 / 60)
    sim.set_neighbor_mode(gpu.NEIGHBORS_KDTREE)
    slots_r, slots_s = [], []
    for i in range(c.n):
        start = c.pos[i - 1] if (i % 50 == 49) else c.pos[i]  # some spawns land on an agent: ValidSpawnLocation refuses
        slots_r.append(ref.spawn(start, c.goal[i], c.radius[i], c.speed[i]))
        slots_s.append(sim.spawn_agent(start, c.goal[i], c.radius[i], c.speed[i]))
    assert slots_r == slots_s and -1 in slots_s
    n = max(slots_s) + 1
    worst = 0.0
    for t in range(120):
        ref.step(1)
        sim.update(1 / 60)
        if t == 40:
            for s in (5, 17, 3):
                ref.destroy_agent(s)
                sim.destroy_agent(s)
            assert ref.spawn(c.pos[5], c.goal[6], 0.3, 1.4) == sim.spawn_agent(c.pos[5], c.goal[6], 0.3, 1.4) == 3
        if t % 20 == 19:
            a, b = sim.state(n), ref.state(n)
            assert np.array_equal(a["active"], b["active"])
            act = b["active"] > 0
            worst = max(worst, float(np.abs(a["pos"][act] - b["pos"][act]).max()))
    a, b = sim.state(n), ref.state(n)
    act = b["active"] > 0
    print(f"drop-in (KD-tree mode) vs the unmodified reference after 120 ticks: max |dp| {worst:.3e}, "
          f"max |dv| {float(np.abs(a['vel'][act] - b['vel'][act]).max()):.3e}")
    assert np.abs(a["vel"][act] - b["vel"][act]).max() <= 5e-3 and worst <= 5e-3
    sim.close()
    ref.close()
