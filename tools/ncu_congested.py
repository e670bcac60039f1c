"""Brings the C3 crowd to its congested state OUTSIDE the profiler's window, then opens the window for a few ticks.

  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_orca|k_fallback|k_attract" \\
      -c 6 -o gpurun_out/<tag>_congested_full python tools/ncu_congested.py [preroll ticks] [profiled ticks]

(Under `ncu` every intercepted launch costs ~0.1 s, 600 ticks of pre-roll 7 minutes; with the window closed they run at
full speed.)  cuProfilerStart / cuProfilerStop of the driver API open and close the window.
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import bench
    from ecmgenerator_b200 import gpu, scenarios as S

    preroll = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    w, c, off, pxy = bench.build_workload(os.environ.get("AB_CONFIG", "c3_1m"), None)
    sim = gpu.GpuSim(w, c.n, float(S.DT), device=0, record_neighbors=False, path_pool_points=int(off[-1]) + 8 * c.n + 4096)
    sim.bulk_load(c.pos, c.radius, c.speed, off, pxy)
    sim.update(preroll)
    sim.sync()
    cu = ctypes.CDLL("libcuda.so.1")
    print("cuProfilerStart", cu.cuProfilerStart(), flush=True)
    sim.update(ticks)
    sim.sync()
    print("cuProfilerStop", cu.cuProfilerStop(), flush=True)
    print("stats", {k: v for k, v in sim.stats().items() if k in ("ticks", "lp3d_runs", "knn_fallbacks")})
    sim.close()


if __name__ == "__main__":
    main()
