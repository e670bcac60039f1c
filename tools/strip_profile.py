"""Per-kernel cost of ONE rank's tick at strip scale, measured on a single GPU.

Runs the C3 crowd as `--strips` in-process strips on cuda:0 (ecmgpu_comm_init_local) for a few ticks.
Under `ncu --metrics gpu__time_duration.sum` the launch list then shows every kernel of a strip tick at
the per-rank problem size (n / strips agents), which is what bounds the multi-GPU tick.
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--strips", type=int, default=8)
    ap.add_argument("--ticks", type=int, default=4)
    ap.add_argument("--config", default="c3_1m")
    args = ap.parse_args()
    import bench
    from ecmgenerator_b200 import multigpu as M

    w, c, off, pxy = bench.build_workload(args.config, None)
    ls = M.LocalStrips(w, c, off, pxy, args.strips, devices=[0], record_neighbors=False)
    ls.update(2)
    ls.sync()
    t0 = time.perf_counter()
    ls.update(args.ticks)
    ls.sync()
    dt = (time.perf_counter() - t0) / args.ticks
    print(f"{args.strips} in-process strips on one GPU: {1e3 * dt:.3f} ms per tick (all strips serialised)")


if __name__ == "__main__":
    main()
