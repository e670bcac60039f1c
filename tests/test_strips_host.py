"""CPU: host-side logic of the strip decomposition (ecmgenerator_b200/multigpu.py), including a
world_size-2 `gloo` run of the ownership / merge protocol the NCCL path uses."""
import os
import socket

import numpy as np
import pytest

from ecmgenerator_b200 import multigpu as M
from ecmgenerator_b200 import scenarios as S


def test_strip_bounds_balance_and_ownership_partition():
    w = S.world_c1()
    c = S.crowd_c1(w, n=4000, seed=4)
    for R in (1, 2, 3, 8):
        b = M.strip_bounds(c.pos[:, 0], R)
        assert len(b) == R + 1 and (np.diff(b) > 0).all()
        own = M.owner_of(c.pos[:, 0], b)
        assert own.min() == 0 and own.max() == R - 1
        counts = np.bincount(own, minlength=R)
        assert counts.max() - counts.min() <= max(8, 0.02 * c.n)
    # ties on the boundary value go to the right-hand strip, like k_pack's `x >= hi`
    x = np.array([0.0, 1.0, 1.0, 1.0, 2.0], np.float32)
    b = M.strip_bounds(x, 2)
    own = M.owner_of(x, b)
    assert (own[x < b[1]] == 0).all() and (own[x >= b[1]] == 1).all()


def test_halo_members_are_exactly_the_agents_within_the_halo():
    rng = np.random.default_rng(1)
    x = rng.uniform(-100, 100, 5000).astype(np.float32)
    b = M.strip_bounds(x, 4)
    own = M.owner_of(x, b)
    for r in range(4):
        idx = M.halo_members(x, b, r, 7.5)
        assert (own[idx] != r).all() and (np.abs(own[idx] - r) == 1).all()
        lo = b[r] if r > 0 else -np.inf
        hi = b[r + 1] if r < 3 else np.inf
        expect = ((x >= lo - 7.5) & (x < lo)) | ((x >= hi) & (x < hi + 7.5))
        assert np.array_equal(np.sort(idx), np.nonzero(expect)[0])


def test_merge_owned_is_bit_exact():
    rng = np.random.default_rng(2)
    a = rng.normal(size=(1000, 2)).astype(np.float32)
    a[::7] = -0.0
    own = rng.integers(0, 3, 1000)
    parts = [np.where((own == r)[:, None], a, np.float32(123.0)) for r in range(3)]
    acc = {}

    def fake_all_reduce(rank):
        def f(bits):
            acc.setdefault("sum", np.zeros_like(bits))
            acc["sum"] += bits
            return acc["sum"]
        return f

    out = None
    for r in range(3):
        out = M.merge_owned(parts[r], (own == r).astype(np.uint8), fake_all_reduce(r))
    assert np.array_equal(out.view(np.uint32), a.view(np.uint32))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = S.world_c1()
        c = S.crowd_c1(w, n=3000, seed=9)
        bounds = M.strip_bounds(c.pos[:, 0], world)
        # every rank derives the same bounds from the same crowd
        t = torch.from_numpy(bounds.copy())
        dist.broadcast(t, src=0)
        same_bounds = bool(np.array_equal(t.numpy(), bounds))
        own = M.owner_of(c.pos[:, 0], bounds)
        mine = (own == rank)
        # each rank holds garbage for the slots it does not own; the merge must reproduce the global array
        local = np.where(mine[:, None], c.pos, np.float32(np.nan))

        def all_reduce_sum(a):
            tt = torch.from_numpy(np.ascontiguousarray(a))
            dist.all_reduce(tt, op=dist.ReduceOp.SUM)
            return tt.numpy()

        merged = M.merge_owned(local, mine.astype(np.uint8), all_reduce_sum)
        owners = all_reduce_sum(mine.astype(np.int64))
        # halo bookkeeping: what I must receive is what my neighbour computes it must send
        need = M.halo_members(c.pos[:, 0], bounds, rank, 9.0)
        box = [None] * world
        dist.all_gather_object(box, need.tolist())
        other = 1 - rank
        x = c.pos[:, 0]
        if rank == 0:
            send_to_other = np.nonzero(mine & (x >= bounds[1] - np.float32(9.0)))[0]
        else:
            send_to_other = np.nonzero(mine & (x < bounds[1] + np.float32(9.0)))[0]
        q.put((rank, same_bounds, bool(np.array_equal(merged.view(np.uint32), c.pos.view(np.uint32))),
               bool((owners == 1).all()), sorted(box[other]) == sorted(send_to_other.tolist())))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_ownership_and_merge():
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same_bounds, merged_ok, owners_ok, halo_ok in res:
        assert same_bounds and merged_ok and owners_ok and halo_ok, (rank, same_bounds, merged_ok, owners_ok, halo_ok)
