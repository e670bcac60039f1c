"""TEST INFRASTRUCTURE: one strip per THREAD of this process, each driving its own handle of the mock-runtime library
through ecmgpu_update like a rank of `torchrun` does on the GPUs - with the NCCL transport (mock_nccl.cpp: mailboxes
between the threads) or the peer transport (the neighbours' inboxes "mapped" through the mock IPC calls, k_exchange_p2p
spinning on the sequence number another thread writes, the tick replayed as a captured graph).  Prints one JSON line.

  python threaded_strips.py <golden scene> <ranks> <nccl|p2p> <compact 0|1> <ticks>
"""
import ctypes
import json
import os
import sys
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)


def main():
    name, n_ranks, transport, compact, ticks = sys.argv[1], int(sys.argv[2]), sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
    import make_mock

    ctypes.CDLL(make_mock.build_nccl(), mode=ctypes.RTLD_GLOBAL)  # the product's dlopen("libnccl.so.2") finds it by name
    os.environ.pop("ECMGPU_COMPACT", None)
    if compact:
        os.environ["ECMGPU_COMPACT"] = "1"
    from ecmgenerator_b200 import gpu
    from ecmgenerator_b200 import multigpu as M
    from tests.util import Golden

    gpu.LIB_PATH, gpu._lib = make_mock.build(), None
    g = Golden(name)
    n = g.n
    p = g.crowd.pos.astype(np.float64)
    r5 = float(np.sqrt(np.sort(((p[:, None, :] - p[None, :, :]) ** 2).sum(-1), axis=1)[:, 5]).max())
    bounds = M.strip_bounds(g.crowd.pos[:, 0], n_ranks)
    widths = np.diff(bounds)[1:-1]
    halo = float(min(2.0 * r5 + 2.0, widths.min())) if len(widths) else 2.0 * r5 + 2.0
    uid = gpu.GpuSim.comm_unique_id()
    barrier = threading.Barrier(n_ranks)
    blobs, out, errors = [None] * n_ranks, [None] * n_ranks, []

    def rank_main(rank):
        try:
            sim = gpu.GpuSim(g.world, n, g.step)
            sim.bulk_load(g.crowd.pos, g.crowd.radius, g.crowd.speed, g.path_off, g.path_xy)
            sim.query_neighbors(1)  # fixes the neighbour cell from the whole crowd, like StripSim
            sim.comm_init(uid, rank, n_ranks)
            sim.comm_set_strips(bounds, halo)
            if transport == "p2p":
                blobs[rank] = sim.comm_p2p_export()
                barrier.wait()
                sim.comm_p2p_connect(blobs[rank - 1] if rank > 0 else None, blobs[rank + 1] if rank < n_ranks - 1 else None)
                barrier.wait()
            for _ in range(ticks):
                sim.update(1)
            sim.sync()
            st = sim.stats()
            out[rank] = {"active": sim.read(gpu.ACTIVE, 0, n), "pos": sim.read(gpu.POS, 0, n), "vel": sim.read(gpu.VEL, 0, n),
                         "halo_misses": st["halo_misses"], "launches": st["kernel_launches"]}
            barrier.wait()  # nobody tears its inboxes down while a neighbour may still write into them
            sim.close()
        except Exception as e:  # noqa: BLE001
            errors.append(f"rank {rank}: {e!r}")
            barrier.abort()

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(n_ranks)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        print(json.dumps({"errors": errors}))
        return 1
    owners = np.stack([o["active"] for o in out]).astype(np.int32).sum(axis=0)
    pos, vel = np.zeros((n, 2), np.float32), np.zeros((n, 2), np.float32)
    for o in out:
        a = o["active"] > 0
        pos[a], vel[a] = o["pos"][a], o["vel"][a]
    mode = "exact-knn"
    gold_act = g.z[f"{mode}/active"][ticks - 1] > 0
    own0 = M.owner_of(g.crowd.pos[:, 0], bounds)
    own1 = M.owner_of(pos[:, 0], bounds)
    print(json.dumps({
        "owners_ok": bool(np.array_equal(owners > 0, gold_act) and owners.max() <= 1),
        "pos_equal": bool(np.array_equal(pos[gold_act].view(np.uint32), g.z[f"{mode}/pos"][ticks - 1][gold_act].view(np.uint32))),
        "vel_equal": bool(np.array_equal(vel[gold_act].view(np.uint32), g.z[f"{mode}/vel"][ticks - 1][gold_act].view(np.uint32))),
        "halo_misses": int(sum(o["halo_misses"] for o in out)), "moved": int(((own0 != own1) & gold_act).sum()),
        "owned": [int((o["active"] > 0).sum()) for o in out]}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
