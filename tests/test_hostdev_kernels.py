"""CPU: the KERNELS of csrc/device/tick.cuh and strips.cuh, compiled for the host and launched as loops with one
thread per block (tests/hostdev/kernels_emul.cpp), in the launch order of ecmgpu_update_phase.

Checks, bit for bit: (1) one simulated device reproduces the reference's golden trajectories - snapshot glue,
LP3D queue and finisher, event lists included; (2) three strips with halo exchange and migration (the in-process
transport's message copy) give the same trajectory and hand every agent to exactly one owner; (3) the record
kernels of ecmgpu_update_io_owned.  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from ecmgenerator_b200 import multigpu as M
from ecmgenerator_b200.gpu import AGENT_REC
from tests.conftest import ROOT
from tests.util import GOLDEN, Golden, assert_bits_equal

HD = os.path.join(ROOT, "tests", "hostdev")
f32p, i32p, u8p, u32p, u64p = (C.POINTER(t) for t in (C.c_float, C.c_int, C.c_uint8, C.c_uint32, C.c_uint64))
C_TOTAL_LP3D, C_TOTAL_HALO_MISS = 6, 9  # enum Counter (csrc/device/tick.cuh)


def _p(a, t):
    return a.ctypes.data_as(t)


@pytest.fixture(scope="session")
def emu():
    return load_emu()


def load_emu(flags=(), tag=""):
    """flags: extra -D switches (e.g. the SIMT emulation); tag names the library file.
    -fno-gnu-unique: several builds of the same device code live in one test process; g++ would otherwise make the
    statics of inline functions (the kernels' `__shared__` arrays) process-wide unique symbols shared by all of them."""
    so = os.path.join(HD, "_build", f"libkernels_emul{tag}.so")
    src = os.path.join(HD, "kernels_emul.cpp")
    dev = os.path.join(ROOT, "ecmgenerator_b200", "csrc", "device")
    deps = [src] + [os.path.join(HD, "shim", f) for f in os.listdir(os.path.join(HD, "shim"))] + [os.path.join(dev, f) for f in os.listdir(dev)]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-gnu-unique", "-fPIC", "-shared", "-I", os.path.join(HD, "shim"),
                               "-o", so, src] + list(flags))
    L = C.CDLL(so)
    vp = C.c_void_p
    L.emu_create.restype = vp
    L.emu_create.argtypes = [C.c_int, f32p, C.c_int, i32p, f32p, C.c_int, f32p, i32p, i32p, u8p, C.c_int, C.c_float, C.c_float, C.c_float,
                             C.c_float, C.c_int, C.c_int]
    L.emu_destroy.argtypes = [vp]
    L.emu_load.argtypes = [vp, C.c_int, f32p, f32p, f32p, i32p, f32p]
    L.emu_set_path.argtypes = [vp, C.c_int, f32p, C.c_int]
    L.emu_destroy_agent.argtypes = [vp, C.c_int]
    L.emu_set_state.argtypes = [vp, f32p, f32p]
    L.emu_set_strips.argtypes = [vp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int]
    L.emu_pack.argtypes = [vp]
    L.emu_exchange.argtypes = [vp, vp, vp]
    L.emu_tick.argtypes = [vp]
    L.emu_tick_kd.argtypes = [vp]
    L.emu_set_compact.argtypes = [vp, C.c_int]
    L.emu_walk_len.argtypes = [vp]
    L.emu_valid_spawn.argtypes = [vp, C.c_int, f32p, f32p, u8p]
    L.emu_kd_reset.argtypes = [vp]
    L.emu_query_neighbors_kd.argtypes = [vp, i32p, i32p]
    L.emu_read.argtypes = [vp, f32p, f32p, f32p, f32p, f32p, u8p, i32p, i32p, u32p, i32p]
    L.emu_counters.argtypes = [vp, u64p]
    L.emu_poll.argtypes = [vp, i32p, i32p, i32p, i32p]
    L.emu_msg_bytes.argtypes = [vp]
    L.emu_get_send.argtypes = [vp, C.c_int, u8p]
    L.emu_set_recv.argtypes = [vp, C.c_int, u8p]
    L.emu_adopt.argtypes = [vp]
    L.emu_collect_owned.argtypes = [vp, vp]
    L.emu_apply_records.argtypes = [vp, C.c_int, vp]
    return L


class EmuDevice:
    def __init__(self, L, g: Golden, cell: float, x_range=None):
        """x_range: the grid covers only this part of the world along x (compact strips, ecmgpu.cu build_grid)."""
        w = g.world
        self.L, self.n = L, g.n
        x0, y0, x1, y1 = (float(v) for v in w.bbox)
        if x_range is not None:
            x0, x1 = max(x0, float(x_range[0])), min(x1, float(x_range[1]))
        gw, gh = int((x1 - x0 + 2 * cell) / cell) + 1, int((y1 - y0 + 2 * cell) / cell) + 1
        keep = [np.ascontiguousarray(a) for a in (w.vert_xy, w.edge_v, w.edge_cl, w.obst_xy, w.obst_next, w.obst_prev, w.obst_convex)]
        self.h = L.emu_create(w.n_vertices, _p(keep[0], f32p), w.n_edges, _p(keep[1], i32p), _p(keep[2], f32p), int(w.obst_next.shape[0]),
                              _p(keep[3], f32p), _p(keep[4], i32p), _p(keep[5], i32p), _p(keep[6], u8p), g.n, np.float32(g.step),
                              np.float32(x0 - cell), np.float32(y0 - cell), np.float32(cell), gw, gh)
        arrs = [np.ascontiguousarray(a, np.float32) for a in (g.crowd.pos, g.crowd.radius, g.crowd.speed, g.path_xy)]
        L.emu_load(self.h, g.n, _p(arrs[0], f32p), _p(arrs[1], f32p), _p(arrs[2], f32p), _p(np.ascontiguousarray(g.path_off, np.int32), i32p),
                   _p(arrs[3], f32p))

    def state(self):
        n = self.n
        out = {"pos": np.zeros((n, 2), np.float32), "vel": np.zeros((n, 2), np.float32), "prefvel": np.zeros((n, 2), np.float32),
               "attraction": np.zeros((n, 2), np.float32), "force": np.zeros((n, 2), np.float32), "active": np.zeros(n, np.uint8),
               "nbr": np.zeros((n, 5), np.int32), "nbr_cnt": np.zeros(n, np.int32), "status": np.zeros(n, np.uint32), "cell": np.zeros(n, np.int32)}
        self.L.emu_read(self.h, _p(out["pos"], f32p), _p(out["vel"], f32p), _p(out["prefvel"], f32p), _p(out["attraction"], f32p),
                        _p(out["force"], f32p), _p(out["active"], u8p), _p(out["nbr"], i32p), _p(out["nbr_cnt"], i32p), _p(out["status"], u32p),
                        _p(out["cell"], i32p))
        return out

    def counters(self):
        c = np.zeros(16, np.uint64)
        self.L.emu_counters(self.h, _p(c, u64p))
        return c

    def poll(self):
        r, d = np.zeros(self.n, np.int32), np.zeros(self.n, np.int32)
        nr, nd = C.c_int(0), C.c_int(0)
        self.L.emu_poll(self.h, _p(r, i32p), C.byref(nr), _p(d, i32p), C.byref(nd))
        return r[: nr.value].copy(), d[: nd.value].copy()

    def destroy_agent(self, slot):
        self.L.emu_destroy_agent(self.h, int(slot))

    def set_state(self, vel, attraction):
        v, a = np.ascontiguousarray(vel, np.float32), np.ascontiguousarray(attraction, np.float32)
        self.L.emu_set_state(self.h, _p(v, f32p), _p(a, f32p))

    def set_path(self, slot, path):
        p = np.ascontiguousarray(path, np.float32).reshape(-1, 2)
        self.L.emu_set_path(self.h, int(slot), _p(p, f32p), len(p))

    def close(self):
        self.L.emu_destroy(self.h)


class EmuStrips:
    """n strips, each an EmuDevice holding all slots and owning its share (the layout of multigpu.LocalStrips)."""

    def __init__(self, L, g: Golden, cell: float, n_strips: int, halo: float, narrow_grid: bool = False):
        self.L, self.n = L, g.n
        self.bounds = M.strip_bounds(g.crowd.pos[:, 0], n_strips)

        def x_range(r):  # what build_grid gives a rank of compact strips: its strip and halo
            if not narrow_grid:
                return None
            return (-np.inf if r == 0 else self.bounds[r] - halo, np.inf if r == n_strips - 1 else self.bounds[r + 1] + halo)

        self.devs = [EmuDevice(L, g, cell, x_range(r)) for r in range(n_strips)]
        for r, d in enumerate(self.devs):
            lo = -np.inf if r == 0 else self.bounds[r]
            hi = np.inf if r == n_strips - 1 else self.bounds[r + 1]
            L.emu_set_strips(d.h, r, n_strips, np.float32(lo), np.float32(hi), np.float32(halo), 512, 128)

    def step(self):
        hs = [d.h for d in self.devs]
        for h in hs:
            self.L.emu_pack(h)
        for r, h in enumerate(hs):
            self.L.emu_exchange(h, hs[r - 1] if r > 0 else None, hs[r + 1] if r + 1 < len(hs) else None)
        return sum(self.L.emu_tick(h) for h in hs)

    def state(self):
        sts = [d.state() for d in self.devs]
        owners = np.stack([s["active"] for s in sts]).astype(np.int32).sum(axis=0)
        out = {k: np.zeros_like(v) for k, v in sts[0].items()}
        for s in sts:
            a = s["active"] > 0
            for k in out:
                out[k][a] = s[k][a]
        out["owners"] = owners
        return out

    def destroy_agent(self, slot):
        for d in self.devs:
            d.destroy_agent(slot)

    def set_path(self, slot, path):
        for d in self.devs:
            d.set_path(slot, path)

    def close(self):
        for d in self.devs:
            d.close()


def _r5_max(g):
    """Largest 5th-neighbour distance in the initial crowd: sizes the grid cell (ring budget) and the halo."""
    p = g.crowd.pos.astype(np.float64)
    d2 = ((p[:, None, :] - p[None, :, :]) ** 2).sum(-1)
    return float(np.sqrt(np.sort(d2, axis=1)[:, 5]).max())


def _cell_for(g):
    return max(2.0, _r5_max(g) / 2.0)  # 8 rings then reach 4 x the largest 5th-neighbour distance


def _run_against_golden(sim, g, step, state, label, mode="exact-knn", max_ticks=None):
    from tests.util import apply_events

    full_at = set(int(t) for t in g.z[f"{mode}/full_at"])
    prev_active = np.ones(g.n, bool)
    for t in range(min(g.ticks(mode), max_ticks or 1 << 30)):
        assert step() == 0, "ring budget exhausted: this harness does not emulate the warp-per-agent fallback"
        st = state()
        gold_act = g.z[f"{mode}/active"][t] > 0
        # an agent destroyed on arrival is not integrated any more: compare what was alive through the tick
        assert np.array_equal(st["active"] > 0, gold_act), f"{label}: active after tick {t}"
        assert_bits_equal(st["pos"][gold_act], g.z[f"{mode}/pos"][t][gold_act], f"{label}: pos after tick {t}")
        assert_bits_equal(st["vel"][gold_act], g.z[f"{mode}/vel"][t][gold_act], f"{label}: vel after tick {t}")
        if t in full_at:
            for k in ("prefvel", "attraction", "force"):
                assert_bits_equal(st[k][gold_act], g.z[f"{mode}/full{t}_{k}"][gold_act], f"{label}: {k} after tick {t}")
        if t == 0:
            assert_bits_equal(st["nbr"][prev_active], g.z[f"{mode}/nbr0_ids"][prev_active], f"{label}: neighbour ids of tick 0")
            assert_bits_equal(st["nbr_cnt"][prev_active], g.z[f"{mode}/nbr0_cnt"][prev_active], f"{label}: neighbour counts of tick 0")
        apply_events(sim, g.events_at(mode, t))
        prev_active = gold_act
    return st


@pytest.mark.parametrize("name", GOLDEN)
def test_kernels_reproduce_reference_trajectories_bitwise(emu, name):
    g = Golden(name)
    d = EmuDevice(emu, g, _cell_for(g))
    _run_against_golden(d, g, lambda: emu.emu_tick(d.h), d.state, name)
    c = d.counters()
    if name == "jam_small":
        assert c[C_TOTAL_LP3D] > 300, "the jam must go through the LP3D queue and its finisher"
    d.close()


@pytest.mark.parametrize("name", ["c2_small", "jam_small"])
def test_three_strips_equal_one_device_bitwise(emu, name):
    g = Golden(name)
    # halo: comfortably more than any agent's 5th-neighbour distance at the start
    r5 = _r5_max(g)
    widths = np.diff(M.strip_bounds(g.crowd.pos[:, 0], 3))[1:-1]
    halo = float(min(2.0 * r5 + 2.0, widths.min()))
    s = EmuStrips(emu, g, _cell_for(g), 3, halo)
    own0 = M.owner_of(g.crowd.pos[:, 0], s.bounds)
    st = _run_against_golden(s, g, s.step, s.state, f"{name} / 3 strips")
    assert st["owners"].max() == 1, "every live agent has exactly one owner"
    misses = sum(int(d.counters()[C_TOTAL_HALO_MISS]) for d in s.devs)
    assert misses == 0, f"halo {halo:.1f} m too small for this scene ({misses} misses)"
    own1 = M.owner_of(st["pos"][:, 0], s.bounds)
    moved = int(((own0 != own1) & (st["active"] > 0)).sum())
    print(f"{name}: halo {halo:.1f} m, {moved} agents changed owner")
    assert moved >= 1, "migration must be exercised"
    s.close()


@pytest.mark.parametrize("name", ["c2_small", "jam_small"])
def test_three_strips_with_compact_walk_equal_one_device_bitwise(emu, name):
    """ECMGPU_COMPACT: pack / cell count / scatter walk a list of the slots a strip may own instead of every slot;
    adopted migrants are appended, nobody is listed twice."""
    g = Golden(name)
    r5 = _r5_max(g)
    widths = np.diff(M.strip_bounds(g.crowd.pos[:, 0], 3))[1:-1]
    halo = float(min(2.0 * r5 + 2.0, widths.min()))
    s = EmuStrips(emu, g, _cell_for(g), 3, halo, narrow_grid=True)
    for dev in s.devs:
        emu.emu_set_compact(dev.h, 1)
    own0 = M.owner_of(g.crowd.pos[:, 0], s.bounds)
    st = _run_against_golden(s, g, s.step, s.state, f"{name} / 3 strips, compact walk")
    assert st["owners"].max() == 1
    assert sum(int(d.counters()[C_TOTAL_HALO_MISS]) for d in s.devs) == 0
    own1 = M.owner_of(st["pos"][:, 0], s.bounds)
    lens = [emu.emu_walk_len(d.h) for d in s.devs]
    owned0 = [int((own0 == r).sum()) for r in range(3)]
    arrived = [int(((own1 == r) & (own0 != r) & (st["active"] > 0)).sum()) for r in range(3)]
    print(f"{name}: list lengths {lens}, owned at the start {owned0}, arrived since {arrived}")
    assert sum(arrived) >= 1, "migration must be exercised"
    for r in range(3):
        assert owned0[r] + arrived[r] <= lens[r] < g.n, "a strip lists its own share plus who came, not the whole crowd"
    # the records of ecmgpu_update_io_owned, collected over the lists: every live agent exactly once, with its state
    seen = np.zeros(g.n, np.int32)
    for d in s.devs:
        rec = np.zeros(g.n, AGENT_REC)
        m = emu.emu_collect_owned(d.h, rec.ctypes.data_as(C.c_void_p))
        r = rec[:m]
        seen[r["slot"]] += 1
        assert_bits_equal(np.stack([r["x"], r["y"]], 1), st["pos"][r["slot"]], "record positions")
    assert np.array_equal(seen, (st["active"] > 0).astype(np.int32))
    s.close()


def test_owned_record_kernels(emu):
    g = Golden("c2_small")
    d = EmuDevice(emu, g, _cell_for(g))
    for _ in range(3):
        emu.emu_tick(d.h)
    d.destroy_agent(7)
    st = d.state()
    rec = np.zeros(g.n, AGENT_REC)
    m = emu.emu_collect_owned(d.h, rec.ctypes.data_as(C.c_void_p))
    assert m == int((st["active"] > 0).sum())
    r = rec[:m]
    assert np.array_equal(np.sort(r["slot"]), np.flatnonzero(st["active"]))
    assert_bits_equal(np.stack([r["x"], r["y"]], 1), st["pos"][r["slot"]], "record positions")
    assert_bits_equal(np.stack([r["vx"], r["vy"]], 1), st["vel"][r["slot"]], "record velocities")
    # write them back shifted, plus a record for the dead slot that must be ignored
    back = r.copy()
    back["x"] += np.float32(1.0)
    extra = np.zeros(1, AGENT_REC)
    extra["slot"], extra["x"] = 7, 123.0
    both = np.concatenate([back, extra])
    emu.emu_apply_records(d.h, len(both), both.ctypes.data_as(C.c_void_p))
    st2 = d.state()
    assert_bits_equal(st2["pos"][r["slot"], 0], st["pos"][r["slot"], 0] + np.float32(1.0), "applied x")
    assert st2["pos"][7, 0] == st["pos"][7, 0]
    d.close()


def test_valid_spawn_kernel_equals_the_reference_scan(emu):
    """k_valid_spawn (grid) against Simulator::ValidSpawnLocation's scan over every agent, restated in float32."""
    g = Golden("c2_small")
    rng = np.random.default_rng(21)
    pos = g.crowd.pos
    for cell in (0.6, 3.0, 40.0):
        d = EmuDevice(emu, g, cell)
        d.destroy_agent(5)  # an inactive agent must not block a spawn
        act = np.ones(g.n, bool)
        act[5] = False
        x0, y0, x1, y1 = (float(v) for v in g.world.bbox)
        q = np.concatenate([
            rng.uniform([x0 - 30, y0 - 30], [x1 + 30, y1 + 30], size=(3000, 2)),          # anywhere, also outside the grid
            pos[rng.integers(0, g.n, 1500)] + rng.normal(0, 0.4, size=(1500, 2)),          # close to agents
            pos[:200] + np.float32([0.5, 0.0]), pos[:200] - np.float32([0.0, 0.25]),       # at (about) the clearance exactly
            pos[5:6],                                                                       # on top of the inactive agent
        ]).astype(np.float32)
        cl = rng.choice(np.float32([0.25, 0.5, 1.0, 7.5]), size=len(q)).astype(np.float32)
        cl[-1] = 0.5
        out = np.zeros(len(q), np.uint8)
        emu.emu_valid_spawn(d.h, len(q), _p(q, f32p), _p(cl, f32p), _p(out, u8p))
        dx = q[:, None, 0] - pos[None, act, 0]
        dy = q[:, None, 1] - pos[None, act, 1]
        d2 = dx * dx + dy * dy  # float32 throughout: fl(fl(dx*dx) + fl(dy*dy))
        assert d2.dtype == np.float32
        want = ~(d2 < (cl * cl)[:, None]).any(axis=1)
        assert np.array_equal(out > 0, want), f"cell {cell}: {int((want != (out > 0)).sum())} of {len(q)} answers differ"
        assert 0.2 < want.mean() < 0.95
        d.close()


def _gloo_strip_worker(rank, world, port, name, q, compact=0):
    """One strip per PROCESS: k_pack fills the fixed-size messages, gloo send / recv carries them to the neighbours
    (what ncclSend / ncclRecv do between the GPUs), k_unpack_migrants adopts, then the tick."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist

    from tests.util import apply_events

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L = load_emu()
        g = Golden(name)
        mode = "exact-knn"
        bounds = M.strip_bounds(g.crowd.pos[:, 0], world)
        widths = np.diff(bounds)
        halo = float(min(2.0 * _r5_max(g) + 2.0, widths.max()))
        d = EmuDevice(L, g, _cell_for(g))
        lo = -np.inf if rank == 0 else bounds[rank]
        hi = np.inf if rank == world - 1 else bounds[rank + 1]
        L.emu_set_strips(d.h, rank, world, np.float32(lo), np.float32(hi), np.float32(halo), 512, 128)
        L.emu_set_compact(d.h, compact)
        nbytes = L.emu_msg_bytes(d.h)
        ok, owned_max = True, 0
        for t in range(g.ticks(mode)):
            L.emu_pack(d.h)
            ops, inbox = [], {}
            for dr, peer in ((0, rank - 1), (1, rank + 1)):
                if 0 <= peer < world:
                    out = np.zeros(nbytes, np.uint8)
                    L.emu_get_send(d.h, dr, _p(out, u8p))
                    inbox[dr] = torch.zeros(nbytes, dtype=torch.uint8)
                    ops += [dist.P2POp(dist.isend, torch.from_numpy(out), peer), dist.P2POp(dist.irecv, inbox[dr], peer)]
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            for dr, buf in inbox.items():
                L.emu_set_recv(d.h, dr, _p(np.ascontiguousarray(buf.numpy()), u8p))
            L.emu_adopt(d.h)
            assert L.emu_tick(d.h) == 0
            st = d.state()
            mine = st["active"] > 0
            owned_max = max(owned_max, int(mine.sum()))
            # every rank checks its own agents against the golden trajectory; ownership is checked globally
            ok &= bool(np.array_equal(st["pos"][mine].view(np.uint32), g.z[f"{mode}/pos"][t][mine].view(np.uint32)))
            ok &= bool(np.array_equal(st["vel"][mine].view(np.uint32), g.z[f"{mode}/vel"][t][mine].view(np.uint32)))
            owners = torch.from_numpy(mine.astype(np.int64))
            dist.all_reduce(owners)
            ok &= bool(np.array_equal(owners.numpy() > 0, g.z[f"{mode}/active"][t] > 0)) and int(owners.max()) <= 1
            apply_events(d, g.events_at(mode, t))
        own0 = M.owner_of(g.crowd.pos[:, 0], bounds) == rank
        arrived = int((mine & ~own0).sum())
        q.put((rank, ok, int(d.counters()[C_TOTAL_HALO_MISS]), arrived, owned_max))
        d.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,compact", [("jam_small", 0), ("jam_small", 1)])
def test_two_process_strips_over_gloo_equal_the_reference_bitwise(emu, name, compact):
    import socket

    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_strip_worker, args=(r, 2, port, name, q, compact)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    print(res)
    for rank, ok, misses, arrived, owned_max in res:
        assert ok, f"rank {rank}: trajectory or ownership differs from the reference"
        assert misses == 0
        assert owned_max < Golden(name).n, "each rank works on its share only"
    assert sum(r[3] for r in res) >= 1, "agents must have migrated between the processes"
