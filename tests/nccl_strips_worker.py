"""torchrun worker of tests/test_gpu_strips.py::test_nccl_strips_match_single_gpu_bitwise."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    from ecmgenerator_b200 import gpu
    from ecmgenerator_b200 import multigpu as M
    from ecmgenerator_b200 import scenarios as S
    from ecmgenerator_b200.host import plan_paths

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, ticks = 20000, 240
    w = S.world_c1()
    c = S.crowd_c1(w, n=n, seed=51)
    off, pxy, ok = plan_paths(w, c.pos, c.goal, c.radius)
    assert ok == n
    strips = M.StripSim(w, c, off, pxy, rank, world, local, record_neighbors=False)
    own0 = M.owner_of(c.pos[:, 0], strips.bounds)
    strips.update(ticks)
    strips.sync()
    pos, owners = strips.gather(gpu.POS)
    vel, _ = strips.gather(gpu.VEL)
    gs = strips.global_stats()
    # 24 more ticks through ecmgpu_update_io_owned (two deep): every rank moves only its share; the union of
    # the records must be the whole crowd and equal the single-GPU state after ticks + 24
    raw = strips.sim
    rout = [gpu.PinnedArray((n,), gpu.AGENT_REC) for _ in range(2)]
    cnt = [gpu.PinnedArray((1,), np.int32) for _ in range(2)]
    last, copied = None, []
    for t in range(24):
        tk = raw.update_io_owned(0, None, rout[t & 1], cnt[t & 1])
        if last is not None:
            raw.io_wait(last)
        last = tk
    raw.io_wait(last)
    m = int(cnt[1].array[0])
    rec = rout[1].array[:m].copy()
    counts = [None] * world
    dist.all_gather_object(counts, m)
    recs = [None] * world
    dist.all_gather_object(recs, rec)
    if rank == 0:
        single = gpu.GpuSim(w, n, float(S.DT), device=local, record_neighbors=False, path_pool_points=int(off[-1] * 1.25) + 4096)
        single.bulk_load(c.pos, c.radius, c.speed, off, pxy)
        single.update(ticks)
        act = single.read(gpu.ACTIVE, 0, n)
        snap_p, snap_v = single.read(gpu.POS, 0, n), single.read(gpu.VEL, 0, n)
        single.update(24)
        act2 = single.read(gpu.ACTIVE, 0, n) > 0
        p2, v2 = single.read(gpu.POS, 0, n), single.read(gpu.VEL, 0, n)
        allrec = np.concatenate(recs)
        o = np.argsort(allrec["slot"])
        sl = allrec["slot"][o]
        io_ok = bool(np.array_equal(sl, np.flatnonzero(act2))
                     and np.array_equal(np.stack([allrec["x"][o], allrec["y"][o]], 1).view(np.uint32), p2[sl].view(np.uint32))
                     and np.array_equal(np.stack([allrec["vx"][o], allrec["vy"][o]], 1).view(np.uint32), v2[sl].view(np.uint32)))
        a = act > 0
        sp, sv = snap_p, snap_v
        own1 = M.owner_of(pos[:, 0], strips.bounds)
        res = {"world": world, "agents": n, "ticks": ticks,
               "pos_equal": bool(np.array_equal(pos[a].view(np.uint32), sp[a].view(np.uint32))),
               "vel_equal": bool(np.array_equal(vel[a].view(np.uint32), sv[a].view(np.uint32))),
               "owners_ok": bool(np.array_equal(owners, act)), "halo_misses": gs["halo_misses"],
               "moved": int(((own0 != own1) & a).sum()), "halo": strips.halo, "p2p": bool(strips.p2p),
               "io_owned_ok": io_ok, "io_owned_counts": [int(x) for x in counts]}
        json.dump(res, open(sys.argv[1], "w"))
    dist.barrier()
    for x in rout + cnt:
        x.free()
    strips.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
