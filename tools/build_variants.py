"""Builds libecmgpu.so variants for A/B runs (tools/ab_variants.py) into variants/ (git-ignored, travels to the GPU box).

  python tools/build_variants.py                      # the standard set
  python tools/build_variants.py name=-DFLAG[,-DFLAG] ...

The kNN switches of the standard set are pinned bit-exact on the CPU by tests/test_hostdev.py: what an A/B decides is
speed.  orca_fast trades bit-exact velocities for SFU arithmetic inside the 1e-4 m/s contract (see below).
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STANDARD = {
    "knn_prune": ["-DECM_KNN_PRUNE"],
    "knn_flat": ["-DECM_KNN_FLAT"],
    "knn_flat_prune": ["-DECM_KNN_FLAT", "-DECM_KNN_PRUNE"],
    "attract_bbox4": ["-DECM_ATTRACT_BBOX4"],
    "knn_twopass": ["-DECM_KNN_TWOPASS"],
    "knn_twopass_prune": ["-DECM_KNN_TWOPASS", "-DECM_KNN_PRUNE"],
    # NOT bit-exact (SFU division / square root / sine in the ORCA half-planes and LP only, geom.cuh): its bar is the
    # 1e-4 m/s per-step velocity tolerance - run `ECMGPU_LIB=variants/libecmgpu_orca_fast.so pytest -m gpu
    # tests/test_gpu_parity.py -k "lockstep or free_running"` next to the A/B and read the printed statistics
    "orca_fast": ["-DECM_ORCA_FAST"],
}


def main():
    import __graft_entry__ as g

    specs = dict(STANDARD)
    if len(sys.argv) > 1:
        specs = {}
        for a in sys.argv[1:]:
            name, _, flags = a.partition("=")
            specs[name] = [f for f in flags.split(",") if f]
    out = os.path.join(ROOT, "variants")
    os.makedirs(out, exist_ok=True)
    src = os.path.join(g.CSRC, "ecmgpu.cu")
    for name, flags in specs.items():
        so = os.path.join(out, f"libecmgpu_{name}.so")
        cmd = [g.NVCC] + g.NVCC_FLAGS + ["-ccbin", g.GXX] + flags + ["-Xptxas", "-v", "-o", so, src, "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            print(r.stderr[-3000:])
            raise SystemExit(f"{name}: build failed")
        lines = r.stderr.splitlines()
        k = [i for i, l in enumerate(lines) if "Compiling entry function" in l and "k_orcaE" in l]
        used = next((l for l in lines[k[0]:k[0] + 6] if "Used" in l), "") if k else ""
        print(f"{name:20s} {' '.join(flags):40s} -> variants/{os.path.basename(so)}   k_orca:{used.split(':')[-1][:60]}")
    print("A/B:  python tools/ab_variants.py base " + " ".join(f"{n}=variants/libecmgpu_{n}.so" for n in specs))


if __name__ == "__main__":
    main()
