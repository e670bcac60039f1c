"""Replays the case saved by tools/nan_probe.py on a small simulator, with one or several library builds."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from ecmgenerator_b200 import gpu
    from ecmgenerator_b200 import scenarios as S

    z = np.load(sys.argv[1])
    libs = sys.argv[2:] or [""]
    w = S.world_c3()
    m = len(z["near"])
    me = int(np.flatnonzero(z["near"] == int(z["agent"]))[0])
    for lib in libs:
        gpu._lib = None
        gpu.LIB_PATH = os.path.abspath(lib) if lib else os.path.join(os.path.dirname(gpu.__file__), "libecmgpu.so")
        sim = gpu.GpuSim(w, m + 8, float(S.DT), device=0, record_neighbors=True, path_pool_points=int(z["path_off"][-1]) + 8 * m + 4096)
        sim.bulk_load(z["pos"], z["radius"], z["speed"], z["path_off"], z["path_xy"])
        sim.write(gpu.VEL, z["vel"])
        sim.write(gpu.ATTRACTION, z["attraction"])
        sim.update(1)
        sim.sync()
        pos, vel = sim.read(gpu.POS, 0, m), sim.read(gpu.VEL, 0, m)
        nbr = sim.read(gpu.NEIGHBORS, 0, m)[me]
        print(json.dumps({"lib": os.path.basename(lib) or "default", "pos": pos[me].tolist(), "vel": vel[me].tolist(), "nbr_global": z["near"][nbr].tolist(),
                          "prefvel": sim.read(gpu.PREFVEL, 0, m)[me].tolist(), "status": int(sim.read(gpu.STATUS, 0, m)[me]),
                          "n_nan": int((~np.isfinite(pos).all(axis=1)).sum())}))
        sim.close()


if __name__ == "__main__":
    main()
