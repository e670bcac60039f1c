"""Resident tick time as the crowd evolves (chunks of 50 ticks), with the LP3D / replan counters per chunk."""
import json
import os
import sys

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
    for chunk in range(12):
        sim.mark(0)
        sim.update(50)
        sim.mark(1)
        sim.sync()
        st = sim.stats()
        out.append({"ticks": 5 + 50 * (chunk + 1), "ms": round(sim.elapsed_ms(0, 1) / 50, 4), "active": st["n_active"],
                    "lp3d_per_tick": (st["lp3d_runs"] - prev["lp3d_runs"]) / 50})
        prev = st
    print(json.dumps(out))


if __name__ == "__main__":
    main()
