"""A/B of build variants of libecmgpu.so on ONE box in ONE process (boxes differ by +-10 %).

  python tools/ab_variants.py name[=path/to/lib.so][,ENV=VALUE...] ...

Builds the workload once, then for every variant loads its library, runs the resident-state timing of
bench.py (5 warm-up ticks, 100 timed ticks as one CUDA-event interval) and the per-phase pass, and
prints one JSON line per variant.  State after the run is checked against the first variant bit for
bit, so a variant that changes results is flagged.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import bench
    from ecmgenerator_b200 import gpu
    from ecmgenerator_b200 import scenarios as S

    specs = sys.argv[1:] or ["base"]
    config = os.environ.get("AB_CONFIG", "c3_1m")
    # AB_PLANNER=device: routes planned on the GPU (the 4 M map: 40 s instead of minutes on the host cores)
    w, c, off, pxy = bench.build_workload(config, int(os.environ["AB_AGENTS"]) if os.environ.get("AB_AGENTS") else None,
                                          planner=os.environ.get("AB_PLANNER", "host"))
    n = c.n
    ref = None
    auto_cell = None  # what build_grid chooses for this crowd (= 1.7 / sqrt(local density) when this was written)
    default_lib = os.environ.get("ECMGPU_LIB") or os.path.join(os.path.dirname(gpu.__file__), "libecmgpu.so")
    for spec in specs:
        parts = spec.split(",")
        name, _, path = parts[0].partition("=")
        env = dict(p.split("=", 1) for p in parts[1:])
        for k in [k for k in os.environ if (k.startswith("ECMGPU_") and k != "ECMGPU_LIB") or k in ("AB_CELL", "AB_CELL_SCALE")]:
            del os.environ[k]
        os.environ.update(env)
        gpu._lib = None
        gpu.LIB_PATH = os.path.abspath(path) if path else default_lib
        # AB_CELL: neighbour-grid cell edge in metres (default: chosen from the crowd's density);
        # AB_CELL_SCALE: the same as a multiple of the automatic choice
        cell = float(os.environ.get("AB_CELL", "0"))
        if os.environ.get("AB_CELL_SCALE"):
            if auto_cell is None:
                probe = gpu.GpuSim(w, n, float(S.DT), device=0, record_neighbors=False, path_pool_points=int(off[-1]) + 8 * n + 4096)
                probe.bulk_load(c.pos, c.radius, c.speed, off, pxy)
                probe.update(1)
                probe.sync()
                auto_cell = probe.stats()["neighbor_cell"]
                probe.close()
            cell = auto_cell * float(os.environ["AB_CELL_SCALE"])
        sim = gpu.GpuSim(w, n, float(S.DT), device=0, record_neighbors=False, path_pool_points=int(off[-1]) + 8 * n + 4096,
                         neighbor_cell=cell)
        sim.bulk_load(c.pos, c.radius, c.speed, off, pxy)
        sim.update(5 + int(os.environ.get("AB_PREROLL", "0")))  # AB_PREROLL: let the crowd congest first (tick cost drifts)
        sim.sync()
        best = 1e9
        for rep in range(3):
            sim.mark(0)
            sim.update(100)
            sim.mark(1)
            sim.sync()
            best = min(best, sim.elapsed_ms(0, 1) / 100)
        sim.set_profiling(True)
        acc = {"grid": 0.0, "attract": 0.0, "orca": 0.0, "tick": 0.0}
        for _ in range(30):
            sim.update(1)
            sim.sync()
            for k, v in sim.last_tick_ms().items():
                acc[k] += v / 30
        sim.set_profiling(False)
        pos = sim.read(gpu.POS, 0, n)
        same, gap = None, None
        if ref is None:
            ref = pos
        else:
            same = bool(np.array_equal(pos.view(np.uint32), ref.view(np.uint32)))
            # variants that trade bit-exactness for speed (ECM_ORCA_FAST): how far apart after the 335 + preroll ticks
            d = (pos.astype(np.float64) - ref.astype(np.float64))
            gap = {"rms_m": float(np.sqrt((d ** 2).sum(axis=1).mean())), "max_m": float(np.abs(d).max()),
                   "rows_bit_identical": float((pos.view(np.uint32) == ref.view(np.uint32)).all(axis=1).mean())}
        print(json.dumps({"variant": name, "env": env, "cell": sim.stats()["neighbor_cell"], "ms_per_tick": round(best, 4), "grid": round(acc["grid"], 4),
                          "attract": round(acc["attract"], 4), "orca": round(acc["orca"], 4), "tick_profiled": round(acc["tick"], 4),
                          "same_state_as_first": same, "position_gap_vs_first": gap}), flush=True)
        sim.close()


if __name__ == "__main__":
    main()
