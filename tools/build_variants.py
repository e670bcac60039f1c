"""Builds libecmgpu.so variants for A/B runs (tools/ab_variants.py) into variants/ (git-ignored, travels to the GPU box).

  python tools/build_variants.py                      # the standard set
  python tools/build_variants.py name=-DFLAG[,-DFLAG] ...

Variants must give the default build's state bit for bit unless they say otherwise; tools/ab_variants.py checks it.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# Round 2 measured the round-1 set on one B200 (profiles/r02a_ab_*.jsonl): orca_fast won and became the default, the
# kNN variants, the split / fused / gather ticks and the four-box prefetch lost and were deleted.  What remains is the
# IEEE build of the ORCA arithmetic, for A/B against the default.
STANDARD = {
    "orca_ieee": ["-DECM_ORCA_IEEE"],
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
