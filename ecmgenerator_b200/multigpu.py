"""Strip decomposition across GPUs (host-side orchestration; the device side is csrc/device/strips.cuh).

The reference has no multi-device path at all; this module is the B200-native
scale-out of its per-tick update.  The world is cut into vertical strips of equal agent count; every
rank holds the static per-agent data of ALL agents (radius, speed, path) and owns the agents whose
x lies in its strip.  Once per tick each rank sends one fixed-size message to each neighbour:
halo agents (position, velocity of agents within `halo` of the border) and migrants (agents that
crossed the border).  Results are bit-identical to the single-GPU run as long as no agent raises
ECMGPU_ST_HALO_MISS (counted in stats()["halo_misses"]).

Two transports:
  * StripSim    - one process per GPU (torchrun), NCCL send/recv issued by libecmgpu itself; torch
                  distributed is only used to hand the NCCL unique id around and to reduce statistics.
  * LocalStrips - all strips inside one process (one or several devices), peer copies; also what
                  the single-GPU test box uses to exercise the exchange logic.
"""
from __future__ import annotations

import os

import numpy as np

from . import gpu
from .scenarios import DT


def strip_bounds(x: np.ndarray, n_ranks: int) -> np.ndarray:
    """n_ranks+1 ascending x boundaries with (almost) equal agent counts per strip.

    bounds[0] / bounds[-1] only bracket the data; the first and last strip extend to infinity."""
    x = np.sort(np.asarray(x, np.float64))
    qs = [x[min(len(x) - 1, (len(x) * r) // n_ranks)] for r in range(1, n_ranks)]
    b = np.array([x[0] - 1.0] + qs + [x[-1] + 1.0], np.float64)
    # strictly ascending even with ties
    for i in range(1, len(b)):
        if b[i] <= b[i - 1]:
            b[i] = np.nextafter(np.float32(b[i - 1]), np.float32(np.inf))
    return b.astype(np.float32)


def owner_of(x: np.ndarray, bounds: np.ndarray) -> np.ndarray:
    """Rank owning each x (first / last strip open-ended), same rule as k_assign_owner / k_pack."""
    inner = np.asarray(bounds[1:-1], np.float32)
    return np.searchsorted(inner, np.asarray(x, np.float32), side="right").astype(np.int32)


def halo_members(x: np.ndarray, bounds: np.ndarray, rank: int, halo: float):
    """Indices a rank must RECEIVE as ghosts: agents of the adjacent strips within `halo` of its borders."""
    x = np.asarray(x, np.float32)
    own = owner_of(x, bounds)
    n_ranks = len(bounds) - 1
    sel = np.zeros(len(x), bool)
    if rank > 0:
        sel |= (own == rank - 1) & (x >= np.float32(bounds[rank]) - np.float32(halo))
    if rank < n_ranks - 1:
        sel |= (own == rank + 1) & (x < np.float32(bounds[rank + 1]) + np.float32(halo))
    return np.nonzero(sel)[0]


def merge_owned(local: np.ndarray, owned: np.ndarray, all_reduce_sum) -> np.ndarray:
    """Global array from per-rank arrays where each slot is valid on exactly one rank.

    Works on the raw bits (int32 views) so that the merge is exact, -0.0 included;
    `all_reduce_sum(int64 ndarray) -> int64 ndarray` is supplied by the transport."""
    a = np.ascontiguousarray(local)
    bits = a.view(np.uint8).reshape(a.shape[0], -1).astype(np.int64)
    bits *= owned.reshape(-1, 1).astype(np.int64)
    out = all_reduce_sum(bits)
    return out.astype(np.uint8).reshape(a.shape[0], -1).view(a.dtype).reshape(a.shape)


# per-slot state that travels with an agent when the borders move (the static components are replicated)
REBALANCE_STATE = (gpu.POS, gpu.VEL, gpu.PREFVEL, gpu.ATTRACTION, gpu.FORCE, gpu.REPLAN_PENDING)


def default_halo(neighbor_cell: float) -> float:
    return 4.0 * float(neighbor_cell)


class LocalStrips:
    """n strips inside one process; strips r lives on devices[r % len(devices)]."""

    def __init__(self, world, crowd, path_off, path_xy, n_strips: int, devices=(0,), halo: float | None = None,
                 neighbor_cell: float = 0.0, record_neighbors: bool = True, step: float = float(DT), bounds=None,
                 rebalance_every: int = 0, rebalance_tolerance: float = 0.15):
        n = crowd.n
        self.n = n
        self.ticks, self.rebalances, self._misses_seen = 0, 0, 0
        self.rebalance_every, self.rebalance_tolerance = int(rebalance_every), float(rebalance_tolerance)
        self.bounds = strip_bounds(crowd.pos[:, 0], n_strips) if bounds is None else np.asarray(bounds, np.float32)
        self.sims = []
        pool = int(path_off[-1] * 1.25) + 4096
        for r in range(n_strips):
            s = gpu.GpuSim(world, n, step, device=devices[r % len(devices)], neighbor_cell=neighbor_cell,
                           record_neighbors=record_neighbors, path_pool_points=pool)
            s.bulk_load(crowd.pos, crowd.radius, crowd.speed, path_off, path_xy)
            self.sims.append(s)
        for s in self.sims:  # fix the (auto) neighbour cell while every strip still sees the whole crowd
            s.query_neighbors(1)
        if halo is None:
            halo = default_halo(self.sims[0].stats()["neighbor_cell"])
        self.halo = float(halo)
        for r, s in enumerate(self.sims):
            s.comm_init_local(r, n_strips, self.sims[r - 1] if r > 0 else None, self.sims[r + 1] if r < n_strips - 1 else None)
        for s in self.sims:
            s.comm_set_strips(self.bounds, self.halo)

    def update(self, ticks: int = 1):
        for _ in range(ticks):
            for phase in (0, 1, 2):
                for s in self.sims:
                    s.update_phase(phase)
            self.ticks += 1
            if self.rebalance_every and self.ticks % self.rebalance_every == 0:
                self.maybe_rebalance()

    def owned_counts(self):
        return [int(s.read(gpu.ACTIVE, 0, self.n).sum()) for s in self.sims]

    def maybe_rebalance(self) -> bool:
        """equal-count borders again once a strip holds more than (1 + tolerance) of its share."""
        self.sync()
        counts = self.owned_counts()
        if max(counts) <= (1.0 + self.rebalance_tolerance) * (sum(counts) / len(counts)):
            return False
        self.rebalance()
        self.rebalances += 1
        return True

    def check_exact(self):
        """Fails loudly if an agent's neighbour search reached beyond the halo since the last check: from that tick on
        the strips no longer compute what one GPU computes.  (Nothing repairs the tick; widen the halo.)"""
        self.sync()
        misses = sum(s["halo_misses"] for s in self.stats())
        if misses > self._misses_seen:
            new, self._misses_seen = misses - self._misses_seen, misses
            raise RuntimeError(f"{new} halo misses: the strips diverged from the single-GPU result; widen the halo (now {self.halo:.2f} m)")

    def sync(self):
        for s in self.sims:
            s.sync()

    def gather(self, which):
        act = [s.read(gpu.ACTIVE, 0, self.n) for s in self.sims]
        owners = np.stack(act).astype(np.int32).sum(axis=0)
        assert owners.max() <= 1, "an agent is owned by two strips"
        out = None
        for s, a in zip(self.sims, act):
            loc = s.read(which, 0, self.n)
            if out is None:
                out = np.zeros_like(loc)
            out[a > 0] = loc[a > 0]
        return out, owners.astype(np.uint8)

    def rebalance(self):
        """New equal-count borders from the current crowd: every strip gets the global state, then re-derives ownership."""
        self.sync()
        state = {w: self.gather(w)[0] for w in REBALANCE_STATE}
        owners = self.gather(gpu.ACTIVE)[1]
        self.bounds = strip_bounds(state[gpu.POS][owners > 0, 0], len(self.sims))
        for s in self.sims:
            for w, a in state.items():
                s.write(w, a)
            s.write(gpu.ACTIVE, owners)
            s.comm_set_strips(self.bounds, self.halo)

    def set_path(self, slot: int, path_xy):
        """A replan answered on EVERY strip: paths are replicated at load time and do not travel with a migrant, so a
        new polyline has to reach all ranks or the agent would follow a stale one after crossing a border."""
        for s in self.sims:
            s.set_path(slot, path_xy)

    def stats(self):
        return [s.stats() for s in self.sims]

    def close(self):
        for s in self.sims:
            s.close()


class StripSim:
    """One strip per process (torchrun): rank r drives cuda:local_rank through libecmgpu + NCCL."""

    def __init__(self, world, crowd, path_off, path_xy, rank: int, n_ranks: int, device: int, halo: float | None = None,
                 neighbor_cell: float = 0.0, record_neighbors: bool = False, step: float = float(DT),
                 rebalance_every: int = 0, rebalance_tolerance: float = 0.15):
        import torch.distributed as dist

        self.rank, self.n_ranks, self.n = rank, n_ranks, crowd.n
        self.ticks, self.rebalances, self._misses_seen = 0, 0, 0
        self.rebalance_every, self.rebalance_tolerance = int(rebalance_every), float(rebalance_tolerance)
        self.sim = gpu.GpuSim(world, crowd.n, step, device=device, neighbor_cell=neighbor_cell, record_neighbors=record_neighbors,
                              path_pool_points=int(path_off[-1] * 1.25) + 4096)
        self.sim.bulk_load(crowd.pos, crowd.radius, crowd.speed, path_off, path_xy)
        self.bounds = strip_bounds(crowd.pos[:, 0], n_ranks)  # same data on every rank -> same bounds
        self.sim.query_neighbors(1)  # fixes the (auto) neighbour cell from the whole crowd, identically on every rank
        if halo is None:
            halo = default_halo(self.sim.stats()["neighbor_cell"])
        self.halo = float(halo)
        box = [gpu.GpuSim.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        self.sim.comm_init(box[0], rank, n_ranks)
        self.sim.comm_set_strips(self.bounds, self.halo)
        # peer transport: halo / migrant entries are stored straight into the neighbours' inboxes over
        # NVLink (CUDA IPC); NCCL stays initialised as the fallback transport (ECMGPU_P2P=0)
        self.p2p = os.environ.get("ECMGPU_P2P", "1") != "0" and n_ranks > 1
        if self.p2p:
            blobs = [None] * n_ranks
            dist.all_gather_object(blobs, self.sim.comm_p2p_export())
            self.sim.comm_p2p_connect(blobs[rank - 1] if rank > 0 else None, blobs[rank + 1] if rank < n_ranks - 1 else None)
            dist.barrier()

    # the bench / tests drive a StripSim like a GpuSim
    def update(self, n: int = 1):
        if not self.rebalance_every:
            self.sim.update(n)
            self.ticks += n
            return
        while n > 0:  # in chunks that end on the re-balancing ticks (a collective: every rank takes the same decisions)
            k = min(n, self.rebalance_every - self.ticks % self.rebalance_every)
            self.sim.update(k)
            self.ticks += k
            n -= k
            if self.ticks % self.rebalance_every == 0:
                self.maybe_rebalance()

    def maybe_rebalance(self) -> bool:
        """equal-count borders again once a strip holds more than (1 + tolerance) of its share
        (collective; the counts travel in one small all-reduce, the state only if the borders do move)."""
        self.sim.sync()
        mine = np.zeros(self.n_ranks, np.int64)
        mine[self.rank] = int(self.sim.read(gpu.ACTIVE, 0, self.n).sum())
        counts = self._all_reduce_sum(mine)
        if counts.max() <= (1.0 + self.rebalance_tolerance) * counts.mean():
            return False
        self.rebalance()
        self.rebalances += 1
        return True

    def check_exact(self):
        """Fails loudly (on every rank) if an agent's neighbour search reached beyond the halo since the last check."""
        misses = self.global_stats(("halo_misses",))["halo_misses"]
        if misses > self._misses_seen:
            new, self._misses_seen = misses - self._misses_seen, misses
            raise RuntimeError(f"{new} halo misses: the strips diverged from the single-GPU result; widen the halo (now {self.halo:.2f} m)")

    def sync(self):
        self.sim.sync()

    def stats(self):
        return self.sim.stats()

    def mark(self, which):
        self.sim.mark(which)

    def elapsed_ms(self, a, b):
        return self.sim.elapsed_ms(a, b)

    def set_profiling(self, on):
        self.sim.set_profiling(on)

    def _all_reduce_sum(self, a: np.ndarray) -> np.ndarray:
        import torch
        import torch.distributed as dist

        t = torch.from_numpy(np.ascontiguousarray(a))
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.cpu().numpy()

    def set_path(self, slot: int, path_xy):
        """A replan answered on this rank's copy; a COLLECTIVE by convention: every rank must call it with the same
        arguments (paths are replicated, they do not travel with a migrant)."""
        self.sim.set_path(slot, path_xy)

    def global_active(self) -> int:
        act = self.sim.read(gpu.ACTIVE, 0, self.n)
        return int(self._all_reduce_sum(np.array([int(act.sum())], np.int64))[0])

    def global_stats(self, keys=("halo_misses", "knn_fallbacks", "lp3d_runs", "location_failures", "replans")) -> dict:
        st = self.sim.stats()
        v = self._all_reduce_sum(np.array([int(st[k]) for k in keys], np.int64))
        return {k: int(x) for k, x in zip(keys, v)}

    def gather(self, which):
        """Global array of `which` (every rank gets it) and the per-slot owner count (must be <= 1)."""
        act = self.sim.read(gpu.ACTIVE, 0, self.n)
        owners = self._all_reduce_sum(act.astype(np.int64))
        loc = self.sim.read(which, 0, self.n)
        return merge_owned(loc, act, self._all_reduce_sum), owners.astype(np.uint8)

    def rebalance(self):
        """New equal-count borders from the current crowd (collective: every rank calls it).  The global state is
        merged through torch.distributed, written to every rank, and ownership re-derived; message buffers,
        peer mappings and sequence numbers stay (ecmgpu_comm_set_strips, re-balancing clause)."""
        import torch.distributed as dist

        self.sim.sync()
        state = {w: self.gather(w)[0] for w in REBALANCE_STATE}
        owners = self.gather(gpu.ACTIVE)[1]
        self.bounds = strip_bounds(state[gpu.POS][owners > 0, 0], self.n_ranks)
        for w, a in state.items():
            self.sim.write(w, a)
        self.sim.write(gpu.ACTIVE, owners)
        self.sim.comm_set_strips(self.bounds, self.halo)
        dist.barrier()

    def close(self):
        self.sim.close()
