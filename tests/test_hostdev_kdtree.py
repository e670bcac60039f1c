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


@pytest.mark.parametrize("take", [1, 2, 3, 5, 6, 7, 40])
def test_kd_mode_tiny_crowds_and_carried_list(emu, take):
    """Fewer agents than list places: the reference's lists then hold ids left over from earlier queries (zeros at
    first, ORCA.h:87) - across agents within a tick and from the last agent of one tick to the first of the next.
    Several ticks against the C oracle in the reference's KD-tree mode, every float."""
    import copy

    from oracle.pyoracle import OracleSim

    g = Golden("c1_small")
    g2 = copy.copy(g)
    g2.crowd = g.crowd.take(np.arange(take))
    g2.n = take
    g2.path_off = g.path_off[: take + 1]
    g2.path_xy = g.path_xy[: g.path_off[take]]
    d = EmuDevice(emu, g2, 6.0)
    emu.emu_kd_reset(d.h)
    ora = OracleSim(g.world, take + 8, g.step, "ref-kdtree")
    ora.bulk_load(g2.crowd.pos, g2.crowd.radius, g2.crowd.speed, g2.path_off, g2.path_xy)
    ids, cnt = _query(emu, d)
    io, co = ora.query_neighbors(take)
    assert_bits_equal(cnt, co, "counts")
    assert_bits_equal(ids, io, "ids")
    for t in range(12):
        emu.emu_tick_kd(d.h)
        ora.step(1)
        a, b = d.state(), ora.state(take)
        for k in ("pos", "vel", "prefvel", "attraction", "force"):
            x, y = a[k][:take], b[k]
            # a lone agent is its own "neighbour" five times over (the zero-filled list names slot 0): the reference
            # divides 0 by 0 (ORCA.cpp:358-362) and gets five NaN constraints - which its LP then passes over, because
            # `if (d <= 0) return i; else if (d > 0) {...}` does neither for a NaN discriminant (ORCA.cpp:499-507), so
            # the agent simply walks at its preferred velocity (checked against oracle/_ref: finite, bit-identical)
            nan = np.isnan(x) & np.isnan(y)
            assert np.array_equal(np.isnan(x), np.isnan(y)), f"{take} agents, tick {t}: {k}: NaN pattern"
            assert_bits_equal(np.where(nan, 0, x).astype(np.float32), np.where(nan, 0, y).astype(np.float32), f"{take} agents, tick {t}: {k}")
        if take == 1 and t == 0:
            assert np.isfinite(a["vel"][0]).all() and np.abs(a["vel"][0]).max() > 0, "the lone agent walks (NaN constraints are passed over)"
        if np.isnan(a["pos"][:take]).any():
            # an agent whose list named itself is NaN now; the reference's next std::sort runs on NaN coordinates
            # (no strict weak order: undefined), so there is nothing left to compare
            break
    print(f"{take} agents: {t + 1} ticks compared")
    d.close()
    ora.close()


def _model_tree(pos):
    """KDTree::Construct (KDTree.cpp:22-83) for a tie-free crowd, in plain Python."""
    import math

    n = len(pos)
    maxd = math.ceil(math.log2(n + 1) - 1)
    tree = [-1] * (2 ** (maxd + 1) - 1)

    def rec(idx, ids, depth):
        if not ids:
            return
        ids = sorted(ids, key=lambda i: pos[i][depth % 2])
        mid = (len(ids) - 1) // 2
        tree[idx] = ids[mid]
        rec(2 * idx + 1, ids[:mid], depth + 1)
        rec(2 * idx + 2, ids[mid + 1:], depth + 1)

    rec(0, list(range(n)), 0)
    return tree, maxd


def _model_query(tree, maxd, pos, a, ids):
    """KDTree::KNearestAgents_R (KDTree.cpp:98-202) writing into the shared list `ids`, in plain Python (float32)."""
    f = np.float32
    eps, dist, found, t = f(1e-4), [f(3.402823466e+38)] * 5, [0], pos[a]

    def rec(cur, depth):
        if depth > maxd or cur >= len(tree) or tree[cur] == -1:
            return
        c = pos[tree[cur]]
        dx, dy = f(c[0] - t[0]), f(c[1] - t[1])
        sq = f(f(dx * dx) + f(dy * dy))
        if found[0] < 5 and sq > eps:
            ids[found[0]] = tree[cur]
            dist[found[0]] = sq
            found[0] += 1
            if found[0] == 5:
                lg, li = sq, 4
                for i in range(4):
                    if dist[i] > lg:
                        lg, li = dist[i], i
                dist[0] = lg; ids[0] = ids[li]; dist[li] = sq; ids[li] = tree[cur]
        else:
            found[0] = 5
            if sq < dist[0] and sq > eps:
                dist[0] = sq; ids[0] = tree[cur]
                lg, li = sq, 0
                for i in range(1, 5):
                    if dist[i] > lg:
                        lg, li = dist[i], i
                dist[0] = lg; ids[0] = ids[li]; dist[li] = sq; ids[li] = tree[cur]
        cv, tv = c[depth % 2], t[depth % 2]
        first, second = (2 * cur + 1, 2 * cur + 2) if tv < cv else (2 * cur + 2, 2 * cur + 1)
        rec(first, depth + 1)
        dd = f(tv - cv)
        if f(dd * dd) < dist[4]:
            rec(second, depth + 1)

    rec(0, 0)
    return found[0]


def test_kd_lists_equal_an_independent_python_model(emu):
    """The lists ORCA used in every tick (ECMGPU_NEIGHBORS) against a plain-Python model of the reference's tree and
    search that keeps ONE list alive across agents and ticks like ORCA::m_NeighborCache."""
    import copy

    g = Golden("c1_small")
    take = 40
    g2 = copy.copy(g)
    g2.crowd, g2.n = g.crowd.take(np.arange(take)), take
    g2.path_off, g2.path_xy = g.path_off[: take + 1], g.path_xy[: g.path_off[take]]
    d = EmuDevice(emu, g2, 6.0)
    emu.emu_kd_reset(d.h)
    carried = [0] * 5
    for t in range(12):
        pos = d.state()["pos"][:take].copy()
        tree, maxd = _model_tree(pos)
        want = []
        for a in range(take):
            _model_query(tree, maxd, pos, a, carried)
            want.append(list(carried))
        emu.emu_tick_kd(d.h)
        got = d.state()["nbr"][:take]
        assert np.array_equal(got, np.asarray(want, np.int32)), f"tick {t}: rows {np.flatnonzero((got != np.asarray(want)).any(1))}"
    d.close()


def test_kd_token_chains_and_the_carried_list(emu):
    """k_kd_resolve / k_kd_cache on hand-made search results.  A search that meets the agent's own node early and is
    pruned soon after leaves list places untouched; the reference then still holds what EARLIER queries wrote there
    (ORCA::m_NeighborCache, ORCA.h:100) - the previous live agent's list, through further untouched places if need be,
    back to the list the previous tick ended with.  In crowds of more than a handful of agents such places are rare
    (the golden scenes have none that survive a search), so the chains are tested directly against a sequential
    restatement: one list, every live agent in slot order writes the places its search wrote."""
    import ctypes as C

    from tests.test_hostdev_kernels import u8p

    emu.emu_kd_resolve.argtypes = [C.c_int, u8p, i32p, i32p, i32p, i32p, i32p]
    rng = np.random.default_rng(4)
    n = 400
    for trial in range(6):
        active = (rng.random(n) < (0.7 if trial else 1.0)).astype(np.uint8)
        if trial == 2:
            active[:37] = 0  # the first live agent is not slot 0
        if trial == 3:
            active[-50:] = 0  # nor is the last live agent the last slot
        raw = rng.integers(0, n, size=(n, 5)).astype(np.int32)
        p_token = [0.0, 0.15, 0.6, 0.97, 0.3, 0.3][trial]  # 0.97: long chains, most of them down to the carried list
        tok = rng.random((n, 5)) < p_token
        place = rng.integers(0, 5, size=(n, 5))  # a token may have been moved: it names the place it came from
        raw[tok] = (-2 - place[tok]).astype(np.int32)
        cnt = rng.integers(0, 6, size=n).astype(np.int32)
        cache = rng.integers(0, n, size=5).astype(np.int32)
        lst, want = list(cache), np.full((n, 5), -9, np.int32)
        for i in range(n):
            if not active[i]:
                continue
            prev = list(lst)
            lst = [int(raw[i, j]) if raw[i, j] >= 0 else prev[-2 - int(raw[i, j])] for j in range(5)]
            want[i] = lst
        nbr, nbr_cnt = np.full((n, 5), -9, np.int32), np.full(n, -9, np.int32)
        cache_io = cache.copy()
        emu.emu_kd_resolve(n, _p(active, u8p), _p(raw, i32p), _p(cnt, i32p), _p(cache_io, i32p), _p(nbr, i32p), _p(nbr_cnt, i32p))
        assert np.array_equal(nbr, want), f"trial {trial}"
        assert np.array_equal(nbr_cnt[active > 0], cnt[active > 0]) and (nbr_cnt[active == 0] == -9).all()
        assert np.array_equal(cache_io, np.asarray(lst, np.int32)), f"trial {trial}: the carried list is the last live agent's"
    # nobody alive: the carried list stays
    cache_io = cache.copy()
    emu.emu_kd_resolve(n, _p(np.zeros(n, np.uint8), u8p), _p(raw, i32p), _p(cnt, i32p), _p(cache_io, i32p), _p(nbr, i32p), _p(nbr_cnt, i32p))
    assert np.array_equal(cache_io, cache)
