"""Finds the first agent whose position turns non-finite, and saves the crowd around it one tick earlier (for a CPU replay)."""
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
    if os.environ.get("PROBE_LIB"):
        gpu._lib = None
        gpu.LIB_PATH = os.path.abspath(os.environ["PROBE_LIB"])
    sim = gpu.GpuSim(w, n, float(S.DT), device=0, record_neighbors=True, path_pool_points=int(off[-1]) + 8 * n + 4096)
    sim.bulk_load(c.pos, c.radius, c.speed, off, pxy)
    start = int(os.environ.get("PROBE_START", "780"))
    sim.update(start)
    sim.sync()
    tag = os.environ.get("PROBE_TAG", "nan")
    for t in range(start, start + int(os.environ.get("PROBE_TICKS", "200"))):
        before = {k: sim.read(getattr(gpu, k.upper()), 0, n) for k in ("pos", "vel", "attraction", "active", "prefvel")}
        sim.update(1)
        sim.sync()
        pos = sim.read(gpu.POS, 0, n)
        act = sim.read(gpu.ACTIVE, 0, n) > 0
        bad = np.flatnonzero(act & ~np.isfinite(pos).all(axis=1))
        if len(bad):
            i = int(bad[0])
            nbr = sim.read(gpu.NEIGHBORS, 0, n)[i]
            p0 = before["pos"][i]
            d = np.linalg.norm(before["pos"] - p0, axis=1)
            near = np.flatnonzero((d < 30.0) & (before["active"] > 0))
            lens = (off[1:] - off[:-1])[near]
            poff = np.zeros(len(near) + 1, np.int32)
            np.cumsum(lens, out=poff[1:])
            idx = np.repeat(off[:-1][near].astype(np.int64), lens) + (np.arange(int(lens.sum())) - np.repeat(poff[:-1].astype(np.int64), lens))
            np.savez_compressed(f"gpurun_out/{tag}_case.npz", tick=t, agent=i, near=near, pos=before["pos"][near], vel=before["vel"][near],
                                attraction=before["attraction"][near], radius=c.radius[near], speed=c.speed[near], path_off=poff, path_xy=pxy[idx],
                                nbr=nbr, after_pos=pos[near], after_vel=sim.read(gpu.VEL, 0, n)[near], status=sim.read(gpu.STATUS, 0, n)[near])
            print(json.dumps({"tick": t, "agent": i, "n_bad": int(len(bad)), "pos_before": p0.tolist(), "vel_before": before["vel"][i].tolist(),
                              "nbr": nbr.tolist(), "nbr_pos": before["pos"][nbr[nbr >= 0]].tolist(), "nbr_vel": before["vel"][nbr[nbr >= 0]].tolist(),
                              "near": int(len(near))}))
            return
    print(json.dumps({"no_nan_until": start + 200}))


if __name__ == "__main__":
    main()
