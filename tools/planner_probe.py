"""Device planner throughput against the number of queries in flight (ECMGPU_PLAN_MB bounds the per-worker scratch).

  python tools/planner_probe.py [queries] [budget_mb ...]
Prints one JSON line per budget: wall time of ecmgpu_plan_paths for `queries` routes of the C3 crowd (after a warm-up
call that allocates the scratch), and whether the polylines equal the first budget's bit for bit.
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from ecmgenerator_b200 import gpu, scenarios as S

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
    budgets = [int(a) for a in sys.argv[2:]] or [0]
    world_fn, crowd_fn = S.CONFIGS[os.environ.get("AB_CONFIG", "c3_1m")]  # starts and goals only: the routes are what is planned here
    w = world_fn()
    c = crowd_fn(w, n=n)
    ref = None
    for mb in budgets:
        if mb:
            os.environ["ECMGPU_PLAN_MB"] = str(mb)
        else:
            os.environ.pop("ECMGPU_PLAN_MB", None)
        sim = gpu.GpuSim(w, 8, float(S.DT))
        t = time.time()
        sim.plan_paths(c.pos[:n], c.goal[:n], c.radius[:n])  # allocates the scratch for n queries in flight
        warm = time.time() - t
        best, kern = 1e9, 1e9
        for _ in range(2):
            t = time.time()
            o, p, ok = sim.plan_paths(c.pos[:n], c.goal[:n], c.radius[:n])
            best = min(best, time.time() - t)
            workers, ms, second = sim.plan_info()
            kern = min(kern, ms)
        same = None
        if ref is None:
            ref = (o.copy(), p.copy())
        else:
            same = bool(np.array_equal(o, ref[0]) and np.array_equal(p.view(np.uint32), ref[1].view(np.uint32)))
        print(json.dumps({"budget_mb": mb, "queries": n, "first_call_s": round(warm, 3), "s": round(best, 3), "us_per_query": round(best / n * 1e6, 3),
                          "workers": workers, "second_pass": second, "kernel_ms": round(kern, 2), "kernel_us_per_query": round(kern * 1e3 / n, 3), "ok": int(ok), "same_as_first": same}), flush=True)
        sim.close()
    if os.environ.get("PROBE_HOST"):  # the host planner (csrc/host/planner.cpp) on all cores of this box, same queries
        from ecmgenerator_b200 import host
        m = min(n, int(os.environ["PROBE_HOST"]))
        t = time.time()
        ho, hp, hok = host.plan_paths(w, c.pos[:m], c.goal[:m], c.radius[:m], threads=0)
        dt = time.time() - t
        print(json.dumps({"host_planner": True, "cores": os.cpu_count(), "queries": m, "s": round(dt, 3), "us_per_query": round(dt / m * 1e6, 3),
                          "same_as_device": bool(m == n and np.array_equal(ho, ref[0]) and np.array_equal(hp.view(np.uint32), ref[1].view(np.uint32)))}), flush=True)


if __name__ == "__main__":
    main()
