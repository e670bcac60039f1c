"""How much of the tick's growth over a long run is the decay of the spatial renumbering?  At tick T the state is read,
loaded into a FRESH simulator (renumbered from the current positions) and both are timed over the same ticks."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import bench
    from ecmgenerator_b200 import gpu
    from ecmgenerator_b200 import scenarios as S

    w, c, off, pxy = bench.build_workload("c3_1m", None)
    n = c.n
    pool = int(off[-1]) + 8 * n + 4096
    sim = gpu.GpuSim(w, n, float(S.DT), device=0, record_neighbors=False, path_pool_points=pool)
    sim.bulk_load(c.pos, c.radius, c.speed, off, pxy)
    for T in (600, 1200):
        sim.update(T - int(sim.stats()["ticks"]))
        sim.sync()
        state = {k: sim.read(getattr(gpu, k), 0, n) for k in ("POS", "VEL", "PREFVEL", "ATTRACTION", "FORCE", "ACTIVE", "REPLAN_PENDING")}
        fresh = gpu.GpuSim(w, n, float(S.DT), device=0, record_neighbors=False, path_pool_points=pool)
        fresh.bulk_load(state["POS"], c.radius, c.speed, off, pxy)  # renumbered from the CURRENT positions
        for k, a in state.items():
            fresh.write(getattr(gpu, k), a)
        res = {}
        for name, s in (("kept", sim), ("renumbered", fresh)):
            s.update(3)
            best = 1e9
            for _ in range(3):
                s.mark(0)
                s.update(50)
                s.mark(1)
                s.sync()
                best = min(best, s.elapsed_ms(0, 1) / 50)
            res[name] = round(best, 4)
        same = bool(np.array_equal(sim.read(gpu.POS, 0, n).view(np.uint32), fresh.read(gpu.POS, 0, n).view(np.uint32)))
        print(json.dumps({"tick": T, "ms_per_tick": res, "same_state": same}), flush=True)
        fresh.close()


if __name__ == "__main__":
    main()
