"""GPU (-m gpu): parity at BASELINE.json's FULL sizes (1 M agents in the C3 city, 4 M in the C4 map, the 250 k corridor
stress) through a size-independent property: a tick is local.  What the device computes for the agents of a window
inside the full crowd must equal what the C oracle computes when it steps only the sub-crowd `window + margin`
(tests/util.py check_window_against_oracle; the property itself is pinned on the CPU by tests/test_window_locality.py):
neighbour lists, ECM cells, attraction points and preferred velocities bit for bit, velocities within 1e-4 m/s.

The oracle cannot step a million agents (its per-agent scans are linear in the map), a window of a few hundred it can.
(The file name sorts last on purpose: these are the longest GPU tests.)"""
import time

import numpy as np
import pytest

from ecmgenerator_b200 import gpu, host
from ecmgenerator_b200 import scenarios as S
from tests.util import check_window_against_oracle

# a kernel that never returns must not hang the box: the watchdog thread ends the run instead
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900, method="thread")]

VEL_TOL = 1e-4  # m/s absolute per step (north_star)


def _paths(w, c, planned):
    """Two-point start -> goal polylines for everybody, the planner's indicative routes for the slots in `planned`."""
    n = c.n
    lens = np.full(n, 2, np.int64)
    poff, pxy, _ = host.plan_paths(w, c.pos[planned], c.goal[planned], c.radius[planned], threads=0)
    plen = np.diff(poff)
    good = plen >= 2
    lens[planned[good]] = plen[good]
    off = np.zeros(n + 1, np.int64)
    np.cumsum(lens, out=off[1:])
    xy = np.zeros((int(off[-1]), 2), np.float32)
    xy[off[:-1]] = c.pos
    xy[off[1:] - 1] = c.goal
    for k in np.flatnonzero(good):
        s = planned[k]
        xy[off[s]:off[s + 1]] = pxy[poff[k]:poff[k + 1]]
    return off.astype(np.int32), xy, int(good.sum())


def _near(pos, windows, reach):
    m = np.zeros(len(pos), bool)
    for x0, y0, x1, y1 in windows:
        m |= (pos[:, 0] >= x0 - reach) & (pos[:, 0] < x1 + reach) & (pos[:, 1] >= y0 - reach) & (pos[:, 1] < y1 + reach)
    return np.flatnonzero(m)


CASES = {
    # config: (window centres, window half size, margin, warm-up ticks)
    "c3_1m": ([(0.0, 0.0), (-1290.0, -1290.0), (610.0, -1300.0)], 30.0, 25.0, 30),
    "c4_4m": ([(0.0, 0.0), (-700.0, 650.0)], 14.0, 12.0, 30),
    "c5_250k": ([(0.0, 0.0), (-300.0, 100.0)], 45.0, 25.0, 60),
}


@pytest.mark.parametrize("config", list(CASES))
def test_full_size_windows_match_the_oracle(config):
    centres, half, margin, warm = CASES[config]
    world_fn, crowd_fn = S.CONFIGS[config]
    w = world_fn()
    c = crowd_fn(w)
    n = c.n
    windows = [(cx - half, cy - half, cx + half, cy + half) for cx, cy in centres]
    t0 = time.time()
    planned = _near(c.pos, windows, margin + 5.0)
    off, pxy, n_planned = _paths(w, c, planned)
    sim = gpu.GpuSim(w, n, float(S.DT), path_pool_points=int(off[-1]) + 8 * n + 4096)
    sim.bulk_load(c.pos, c.radius, c.speed, off, pxy)
    sim.update(warm)
    sim.poll_events()
    before = sim.state(n)
    sim.update(1)
    after = sim.state(n)
    after["nbr"] = sim.read(gpu.NEIGHBORS, 0, n)
    after["nbr_cnt"] = sim.read(gpu.NEIGHBOR_COUNT, 0, n)
    after["cell"] = sim.read(gpu.CELL, 0, n)
    st = sim.stats()
    print(f"{config}: {n} agents ({n_planned} with planned routes), set-up + {warm + 1} ticks in {time.time() - t0:.1f} s; "
          f"lp3d {st['lp3d_runs']}, knn fallbacks {st['knn_fallbacks']}, location failures {st['location_failures']}")
    assert st["n_active"] >= 0.99 * n
    for win in windows:
        t1 = time.time()
        res = check_window_against_oracle(w, float(S.DT), before, after, c.radius, c.speed, off, pxy, win, margin, VEL_TOL,
                                          label=f"{config} window {win}")
        print(f"{config} window {win}: {res} ({time.time() - t1:.1f} s of oracle)")
        assert res["moving"] > 0.5, "the window's agents must be under way (non-trivial ORCA input)"
        assert res["velocity_rows_bit_identical"] > 0.75  # SFU arithmetic in ORCA: see tests/test_gpu_parity.py
    sim.close()
