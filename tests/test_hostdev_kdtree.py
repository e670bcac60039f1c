"""CPU: the faithful KD-tree neighbour mode (csrc/device/kdtree.cuh) compiled for the host and
launched in the order of ecmgpu.cu's enqueue_kd_orca (tests/hostdev/kernels_emul.cpp; std::stable_sort stands in
for the per-level radix sort).

The golden files hold trajectories of the UNMODIFIED reference ("ref-kdtree": its own KDTree.cpp, over-pruning,
duplicated and stale ids included).  Checked bit for bit: every position and velocity of every tick, the full
component arrays at the recorded ticks, and the neighbour lists id for id.  Test infrastructure only."""
import numpy as np
import pytest

from tests.test_hostdev_kernels import EmuDevice, _cell_for, _p, _run_against_golden, emu, i32p  # noqa: F401
from tests.util import GOLDEN, Golden, assert_bits_equal

C_TOTAL_KD_TIES = 10  # enum Counter (csrc/device/tick.cuh)


class _Tie(Exception):
    pass


def _query(emu, d):
    ids = np.zeros((d.n, 5), np.int32)
    cnt = np.zeros(d.n, np.int32)
    emu.emu_query_neighbors_kd(d.h, _p(ids, i32p), _p(cnt, i32p))
    return ids, cnt


@pytest.mark.parametrize("name", GOLDEN)
def test_kd_mode_reproduces_the_unmodified_reference_bitwise(emu, name):
    g = Golden(name)
    mode = "ref-kdtree"
    d = EmuDevice(emu, g, _cell_for(g))
    emu.emu_kd_reset(d.h)
    # the reference's lists on the initial crowd (Simulator::FindNNearestNeighbors with one shared vector)
    ids, cnt = _query(emu, d)
    assert_bits_equal(cnt, g.z[f"{mode}/nbr0_cnt"], f"{name}: neighbour counts before tick 0")
    assert_bits_equal(ids, g.z[f"{mode}/nbr0_ids"], f"{name}: neighbour ids before tick 0")
    differ = (np.sort(ids, 1) != np.sort(g.z["exact-knn/nbr0_ids"], 1)).any(1).mean()
    print(f"{name}: {100 * differ:.1f} % of the reference's lists are not the exact 5-NN")

    # A tick whose tree has a segment with a tie at the median is std::sort-defined in the reference itself: the
    # kernels report it (kd_median_ties) and the comparison ends there.  Up to that tick everything is bit-exact.
    ticks = g.ticks(mode)

    stepped = [0]

    def step():
        emu.emu_tick_kd(d.h)
        if int(d.counters()[C_TOTAL_KD_TIES]) > 0:
            raise _Tie()
        stepped[0] += 1
        return 0

    done = ticks
    try:
        st = _run_against_golden(d, g, step, d.state, f"{name} / kd", mode=mode)
    except _Tie:
        done = stepped[0]
    print(f"{name}: {done} of {ticks} ticks compared bit for bit" + ("" if done == ticks else " (then a median tie: the reference's tree is std::sort-defined)"))
    assert done >= min(ticks, 96), "the golden scenes run tie-free for at least 96 ticks"
    if done == ticks:
        live = st["active"] > 0
        ids, cnt = _query(emu, d)
        assert_bits_equal(cnt[live], g.z[f"{mode}/nbr1_cnt"][live], f"{name}: neighbour counts after the run")
        assert_bits_equal(ids[live], g.z[f"{mode}/nbr1_ids"][live], f"{name}: neighbour ids after the run")
    d.close()


def test_kd_mode_differs_from_exact_mode():
    """The two golden trajectories are different (otherwise the test above would prove nothing about the quirks)."""
    g = Golden("c2_small")
    assert not np.array_equal(g.z["ref-kdtree/final_pos"], g.z["exact-knn/final_pos"])


def test_kd_small_segment_ties_resolve_like_the_reference():
    """Agents with EQUAL coordinates: where the tie sits at the median of a segment of up to 16 agents, libstdc++'s
    std::sort is an insertion sort (stable) and the stable radix sort of the build resolves it the same way - the
    lists must still equal the reference's (C oracle with its std::sort restatement, and the compiled reference
    itself where it is available).  Ties at the median of larger segments are the ones reported as ambiguous."""
    import copy

    from oracle import pyref
    from oracle.pyoracle import OracleSim
    from tests.test_hostdev_kernels import load_emu

    L = load_emu()
    g = Golden("c2_small")
    rng = np.random.default_rng(11)
    pos = g.crowd.pos.copy()
    d2 = ((pos[:, None, :] - pos[None, :, :]) ** 2).sum(-1)
    np.fill_diagonal(d2, np.inf)

    def probe(p):
        g2 = copy.copy(g)
        g2.crowd = copy.copy(g.crowd)
        g2.crowd.pos = p
        dev = EmuDevice(L, g2, _cell_for(g))
        L.emu_kd_reset(dev.h)
        ids, cnt = _query(L, dev)
        c = dev.counters()
        dev.close()
        return ids, cnt, int(c[C_TOTAL_KD_TIES + 1]), int(c[C_TOTAL_KD_TIES])

    # tie nearest neighbours on x or y, one pair at a time; keep a pair unless it puts a tie on the median of a LARGE segment
    used = np.zeros(g.n, bool)
    pairs = 0
    for a in rng.permutation(g.n):
        b = int(np.argmin(d2[a]))
        if used[a] or used[b]:
            continue
        trial = pos.copy()
        trial[b, pairs & 1] = trial[a, pairs & 1]
        if probe(trial)[3] > 0:
            continue
        pos = trial
        used[a] = used[b] = True
        pairs += 1
        if pairs == 60:
            break
    ids, cnt, small, big = probe(pos)
    print(f"{pairs} tied pairs: {small} median ties in segments of up to 16 agents, {big} in larger ones")
    assert small >= 5, "the scene must put ties on medians of small segments"
    assert big == 0
    ora = OracleSim(g.world, g.n + 8, g.step, "ref-kdtree")
    ora.bulk_load(pos, g.crowd.radius, g.crowd.speed, g.path_off, g.path_xy)
    io, co = ora.query_neighbors(g.n)
    assert_bits_equal(cnt, co, "counts vs the C oracle")
    assert_bits_equal(ids, io, "ids vs the C oracle (std::sort restated)")
    if pyref.available("ref-kdtree"):
        r = pyref.RefSim(g.world, g.n + 8, g.step, "ref-kdtree")
        r.bulk_load_paths(pos, g.crowd.radius, g.crowd.speed, g.path_off, g.path_xy) if hasattr(r, "bulk_load_paths") else \
            r.bulk_load(pos, g.crowd.goal, g.crowd.radius, g.crowd.speed)
        ir, cr = r.query_neighbors(g.n)
        r.close()
        assert_bits_equal(ids, ir, "ids vs the compiled reference")
