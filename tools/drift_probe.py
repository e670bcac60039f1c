"""Resident tick time as the crowd evolves (chunks of 50 ticks), with the LP3D / replan counters per chunk."""
import json
import os
import sys

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
    out = []
    prev = sim.stats()
    chunks = int(os.environ.get("DRIFT_CHUNKS", "24"))
    for chunk in range(chunks):
        sim.mark(0)
        sim.update(50)
        sim.mark(1)
        sim.sync()
        st = sim.stats()
        row = {"ticks": 5 + 50 * (chunk + 1), "ms": round(sim.elapsed_ms(0, 1) / 50, 4), "active": st["n_active"],
               "lp3d_per_tick": (st["lp3d_runs"] - prev["lp3d_runs"]) / 50, "replans": st["replans"] - prev["replans"],
               "location_failures": st["location_failures"] - prev["location_failures"], "knn_fallbacks": st["knn_fallbacks"] - prev["knn_fallbacks"]}
        prev = st
        # one profiled tick + the state of the crowd: densest neighbour cell, agents at rest, agents far from any street
        sim.set_profiling(True)
        sim.update(1)
        sim.sync()
        row["phase_ms"] = {k: round(v, 4) for k, v in sim.last_tick_ms().items()}
        sim.set_profiling(False)
        pos, vel, act = sim.read(gpu.POS, 0, n), sim.read(gpu.VEL, 0, n), sim.read(gpu.ACTIVE, 0, n) > 0
        cell = st["neighbor_cell"]
        k = np.floor((pos[act] - pos[act].min(axis=0)) / cell).astype(np.int64)
        occ = np.bincount(k[:, 1] * (k[:, 0].max() + 1) + k[:, 0])
        spd = np.linalg.norm(vel[act], axis=1)
        row.update(max_cell=int(occ.max()), cells_over_40=int((occ > 40).sum()), mean_speed=round(float(spd.mean()), 4), at_rest=int((spd < 0.05).sum()),
                   nan_pos=int(np.isnan(pos[act]).any(axis=1).sum()))
        out.append(row)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
