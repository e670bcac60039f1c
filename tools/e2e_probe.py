"""Where does the end-to-end tick (ecmgpu_update_io) spend its time beyond the resident tick?

Times the pipelined host loop of bench.py with parts of the traffic switched off:
resident ticks | update_io without buffers | uploads only | downloads only | both.
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import bench
    from ecmgenerator_b200 import gpu
    from ecmgenerator_b200 import scenarios as S

    w, c, off, pxy = bench.build_workload(os.environ.get("AB_CONFIG", "c3_1m"), None)
    n = c.n
    sim = gpu.GpuSim(w, n, float(S.DT), device=0, record_neighbors=False, path_pool_points=int(off[-1]) + 8 * n + 4096)
    sim.bulk_load(c.pos, c.radius, c.speed, off, pxy)
    sim.update(5)
    sim.sync()
    D, K = 3, 60
    hp = [gpu.PinnedArray((n, 2), np.float32) for _ in range(D)]
    hv = [gpu.PinnedArray((n, 2), np.float32) for _ in range(D)]
    op = [gpu.PinnedArray((n, 2), np.float32) for _ in range(D)]
    ov = [gpu.PinnedArray((n, 2), np.float32) for _ in range(D)]
    oa = [gpu.PinnedArray((n,), np.uint8) for _ in range(D)]
    p0, v0 = sim.read(gpu.POS, 0, n), sim.read(gpu.VEL, 0, n)
    for g in range(D):
        hp[g].array[:] = p0
        hv[g].array[:] = v0

    def loop(up, down, consume, depth=D, steps=K):
        tickets = []
        for i in range(steps):
            g = i % D
            tickets.append(sim.update_io(n, hp[g] if up else None, hv[g] if up else None, op[g] if down else None,
                                         ov[g] if down else None, oa[g] if down else None))
            j = i - (depth - 1)
            if j >= 0:
                sim.io_wait(tickets[j])
                if consume:
                    int(np.count_nonzero(oa[j % D].array)) if consume == 1 else int((oa[j % D].array > 0).sum())
        for j in range(max(0, steps - (depth - 1)), steps):
            sim.io_wait(tickets[j])

    def timed(fn):
        fn()
        sim.sync()
        t0 = time.perf_counter()
        fn()
        sim.sync()
        return 1e3 * (time.perf_counter() - t0) / K

    res = {"resident_update": timed(lambda: sim.update(K))}
    res["io_no_buffers"] = timed(lambda: loop(False, False, False))
    res["io_up_only"] = timed(lambda: loop(True, False, False))
    res["io_down_only"] = timed(lambda: loop(False, True, False))
    res["io_both"] = timed(lambda: loop(True, True, False))
    res["io_both_consume"] = timed(lambda: loop(True, True, 1))            # np.count_nonzero over the 1 M flags
    res["io_both_consume_slow_numpy"] = timed(lambda: loop(True, True, 2))  # (flags > 0).sum(): what bench.py did until r03z
    res["io_both_depth2"] = timed(lambda: loop(True, True, 1, depth=2))
    os.environ["X"] = "1"
    print(json.dumps({k: round(v, 4) for k, v in res.items()}))


if __name__ == "__main__":
    main()
