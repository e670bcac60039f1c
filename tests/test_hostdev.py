"""CPU: the DEVICE algorithms themselves (csrc/device/{locate,knn,orca}.cuh), compiled for the host with a shim
(tests/hostdev) and run one agent at a time, against the reference's golden vectors - bit for bit.

The GPU tests hold the CUDA build to 1e-4 m/s because CUDA's sinf / cosf / atanf round differently from the C
library's; here the same source runs with the C library's functions, so every float of every tick must equal
the unmodified reference's.  Test infrastructure only: nothing in the product can reach this build."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests.conftest import ROOT
from tests.util import GOLDEN, Golden, assert_bits_equal

HD = os.path.join(ROOT, "tests", "hostdev")
f32p, i32p, u8p, u32p = (C.POINTER(t) for t in (C.c_float, C.c_int, C.c_uint8, C.c_uint32))


def _p(a, t):
    return a.ctypes.data_as(t)


# builds of the device code that must be bit-exact: the plain per-lane instantiations and
VARIANTS = {"default": [],
            # the warp-synchronous instantiations the kernels run (flattened control flow), one lane per "warp"
            "sync": ["-DHD_SYNC"]}


@pytest.fixture(scope="session", params=list(VARIANTS))
def hd(request):
    so = os.path.join(HD, "_build", f"libhostdev_{request.param}.so")
    src = os.path.join(HD, "hostdev.cpp")
    dev = os.path.join(ROOT, "ecmgenerator_b200", "csrc", "device")
    deps = [src] + [os.path.join(HD, "shim", f) for f in os.listdir(os.path.join(HD, "shim"))] + \
        [os.path.join(dev, f) for f in os.listdir(dev)]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        # -ffp-contract=off == nvcc -fmad=false: IEEE mul / add without contraction; div / sqrt are IEEE on both sides
        subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-gnu-unique", "-fPIC", "-shared", "-I", os.path.join(HD, "shim"),
                               "-o", so, src] + VARIANTS[request.param])
    L = C.CDLL(so)
    L.hd_world.restype = C.c_void_p
    L.hd_world.argtypes = [C.c_int, f32p, C.c_int, i32p, f32p, C.c_int, f32p, i32p, i32p, u8p]
    L.hd_world_free.argtypes = [C.c_void_p]
    L.hd_locate.argtypes = [C.c_void_p, C.c_int, f32p, i32p]
    L.hd_retract.argtypes = [C.c_void_p, C.c_int, f32p, u8p, f32p, i32p]
    L.hd_tick.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_int, f32p, f32p, f32p, f32p, f32p, f32p, f32p, u8p, i32p, f32p,
                          i32p, i32p, i32p, i32p, i32p, i32p, u32p]
    L.hd_neighbors.argtypes = [C.c_int, C.c_float, C.c_int, f32p, u8p, i32p, i32p]
    L.hd_considered.restype = C.c_longlong
    L.hd_considered.argtypes = [C.c_int]
    L.variant = request.param
    return L


class HostDevSim:
    """Slot arrays on the host + hd_tick; the interface of OracleSim as far as the golden protocol needs it."""

    def __init__(self, L, g: Golden, cell: float, max_ring: int = 8):
        w = g.world
        self.L, self.n, self.step_s, self.cell, self.max_ring = L, g.n, np.float32(g.step), np.float32(cell), max_ring
        keep = [np.ascontiguousarray(a) for a in (w.vert_xy, w.edge_v, w.edge_cl, w.obst_xy, w.obst_next, w.obst_prev, w.obst_convex)]
        self.h = L.hd_world(w.n_vertices, _p(keep[0], f32p), w.n_edges, _p(keep[1], i32p), _p(keep[2], f32p), int(w.obst_next.shape[0]),
                            _p(keep[3], f32p), _p(keep[4], i32p), _p(keep[5], i32p), _p(keep[6], u8p))
        n = self.n
        self.pos = g.crowd.pos.astype(np.float32).copy()
        self.vel = np.zeros((n, 2), np.float32)
        self.pref = np.zeros((n, 2), np.float32)
        self.attr = np.zeros((n, 2), np.float32)
        self.force = np.zeros((n, 2), np.float32)
        self.radius = g.crowd.radius.astype(np.float32).copy()
        self.speed = g.crowd.speed.astype(np.float32).copy()
        self.active = np.ones(n, np.uint8)
        self.paths = [g.path_xy[g.path_off[i]:g.path_off[i + 1]].astype(np.float32) for i in range(n)]
        self.status = np.zeros(n, np.uint32)
        self.nbr = np.full((n, 5), -1, np.int32)
        self.nbr_cnt = np.zeros(n, np.int32)
        self.fallbacks = 0
        self.lp3d = 0

    def step(self):
        n = self.n
        off = np.zeros(n + 1, np.int32)
        np.cumsum([len(p) for p in self.paths], out=off[1:])
        pxy = np.ascontiguousarray(np.concatenate(self.paths), np.float32)
        rep, des = np.zeros(n, np.int32), np.zeros(n, np.int32)
        nr, nd = C.c_int(0), C.c_int(0)
        self.fallbacks += self.L.hd_tick(self.h, n, self.step_s, self.cell, self.max_ring, _p(self.pos, f32p), _p(self.vel, f32p),
                                         _p(self.pref, f32p), _p(self.attr, f32p), _p(self.force, f32p), _p(self.radius, f32p),
                                         _p(self.speed, f32p), _p(self.active, u8p), _p(off, i32p), _p(pxy, f32p), _p(self.nbr, i32p),
                                         _p(self.nbr_cnt, i32p), _p(rep, i32p), C.byref(nr), _p(des, i32p), C.byref(nd), _p(self.status, u32p))
        self.lp3d += int(((self.status & 64) != 0).sum())
        return rep[: nr.value].copy(), des[: nd.value].copy()

    def neighbors(self):
        ids, cnt = np.full((self.n, 5), -1, np.int32), np.zeros(self.n, np.int32)
        self.L.hd_neighbors(self.n, self.cell, self.max_ring, _p(self.pos, f32p), _p(self.active, u8p), _p(ids, i32p), _p(cnt, i32p))
        return ids, cnt

    # apply_events protocol (tests/util.py)
    def destroy_agent(self, slot):
        self.active[slot] = 0

    def set_path(self, slot, path):
        self.paths[slot] = np.ascontiguousarray(path, np.float32).reshape(-1, 2)

    def close(self):
        self.L.hd_world_free(self.h)


@pytest.mark.parametrize("name", GOLDEN)
@pytest.mark.parametrize("cell", [1.3, 4.0])
def test_device_algorithms_reproduce_reference_trajectories_bitwise(hd, name, cell):
    from tests.util import apply_events

    g = Golden(name)
    mode = "exact-knn"
    s = HostDevSim(hd, g, cell)
    ids, cnt = s.neighbors()
    assert_bits_equal(ids, g.z[f"{mode}/nbr0_ids"], "neighbour ids at t=0")
    assert_bits_equal(cnt, g.z[f"{mode}/nbr0_cnt"], "neighbour counts at t=0")
    full_at = set(int(t) for t in g.z[f"{mode}/full_at"])
    for t in range(g.ticks(mode)):
        s.step()
        assert_bits_equal(s.pos, g.z[f"{mode}/pos"][t], f"pos after tick {t}")
        assert_bits_equal(s.vel, g.z[f"{mode}/vel"][t], f"vel after tick {t}")
        assert np.array_equal(s.active, g.z[f"{mode}/active"][t]), f"active after tick {t}"
        if t in full_at:
            a = g.z[f"{mode}/active"][t] > 0
            assert_bits_equal(s.pref[a], g.z[f"{mode}/full{t}_prefvel"][a], f"prefvel after tick {t}")
            assert_bits_equal(s.attr[a], g.z[f"{mode}/full{t}_attraction"][a], f"attraction after tick {t}")
            assert_bits_equal(s.force[a], g.z[f"{mode}/full{t}_force"][a], f"force after tick {t}")
        apply_events(s, g.events_at(mode, t))
    ids, cnt = s.neighbors()
    assert_bits_equal(ids, g.z[f"{mode}/nbr1_ids"], "neighbour ids at the end")
    assert_bits_equal(cnt, g.z[f"{mode}/nbr1_cnt"], "neighbour counts at the end")
    if name == "jam_small":
        assert s.lp3d > 300, "the jam must exercise RandomizedLP3D"
    s.close()


@pytest.mark.parametrize("name", GOLDEN)
def test_device_point_location_and_retraction(hd, name):
    g = Golden(name)
    s = HostDevSim(hd, g, 2.0)
    pts = np.ascontiguousarray(g.z["probe/xy"], np.float32)
    n = len(pts)
    cells = np.zeros(n, np.int32)
    hd.hd_locate(s.h, n, _p(pts, f32p), _p(cells, i32p))
    assert np.array_equal(cells, g.z["probe/cell"])
    ok, xy, edge = np.zeros(n, np.uint8), np.zeros((n, 2), np.float32), np.zeros(n, np.int32)
    hd.hd_retract(s.h, n, _p(pts, f32p), _p(ok, u8p), _p(xy, f32p), _p(edge, i32p))
    assert np.array_equal(ok, g.z["probe/retract_ok"])
    good = ok > 0
    assert np.array_equal(edge[good], g.z["probe/retract_edge"][good])
    assert_bits_equal(xy[good], g.z["probe/retract_xy"][good], "retracted points")
    s.close()


def _brute_knn(pos, active):
    """The neighbour contract in numpy: float32 sqDist = fl(fl(dx*dx) + fl(dy*dy)), keep > 1e-4, 5 smallest by (sqDist, slot)."""
    n = len(pos)
    ids, cnt = np.full((n, 5), -1, np.int32), np.zeros(n, np.int32)
    act = np.flatnonzero(active)
    for i in act:
        dx = (pos[act, 0] - pos[i, 0]).astype(np.float32)
        dy = (pos[act, 1] - pos[i, 1]).astype(np.float32)
        d = (dx * dx).astype(np.float32) + (dy * dy).astype(np.float32)
        keep = d > np.float32(1e-4)
        order = np.lexsort((act[keep], d[keep]))[:5]
        ids[i, : len(order)] = act[keep][order]
        cnt[i] = len(order)
    return ids, cnt


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_device_knn_equals_brute_force_on_hostile_inputs(hd, seed):
    """Clusters, exact ties (points on an integer lattice), duplicates, sparse outliers far outside the bulk, inactive
    slots - for several grid cell sizes, including cells far smaller and far larger than the spacing."""
    rng = np.random.default_rng(seed)
    parts = [rng.integers(-6, 7, size=(300, 2)).astype(np.float32),                    # lattice: many exact ties and duplicates
             rng.normal(0, 0.4, size=(200, 2)).astype(np.float32) + np.float32(20.0),    # a tight cluster
             rng.uniform(-40, 40, size=(200, 2)).astype(np.float32),                     # background
             np.array([[500.0, -300.0], [-800.0, 10.0], [0.0, 900.0]], np.float32)]       # far outliers: many empty rings
    pos = np.ascontiguousarray(np.concatenate(parts))
    n = len(pos)
    active = (rng.uniform(size=n) > 0.1).astype(np.uint8)
    want_ids, want_cnt = _brute_knn(pos, active)
    for cell in (0.3, 1.0, 2.5, 9.0, 150.0):
        ids, cnt = np.full((n, 5), -1, np.int32), np.zeros(n, np.int32)
        hd.hd_neighbors(n, np.float32(cell), 8, _p(pos, f32p), _p(active, u8p), _p(ids, i32p), _p(cnt, i32p))
        a = active > 0
        assert np.array_equal(cnt[a], want_cnt[a]), f"counts, cell {cell}"
        assert np.array_equal(ids[a], want_ids[a]), f"ids, cell {cell}"


def test_report_candidates_visited(hd):
    """Not a check: prints how many neighbour candidates the variant visits on the dense golden crowd."""
    g = Golden("c2_small")
    s = HostDevSim(hd, g, 1.7 / np.sqrt(g.n / 6000.0))
    hd.hd_considered(1)
    s.neighbors()
    print(f"[{hd.variant}] candidates per agent: {hd.hd_considered(1) / g.n:.1f}")
    s.close()
