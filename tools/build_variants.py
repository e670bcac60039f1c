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
    import shutil
    import tempfile

    import __graft_entry__ as g

    specs = dict(STANDARD)
    if len(sys.argv) > 1:
        specs = {}
        for a in sys.argv[1:]:
            name, _, flags = a.partition("=")
            specs[name] = [f for f in flags.split(",") if f]
    out = os.path.join(ROOT, "variants")
    os.makedirs(out, exist_ok=True)
    for name, flags in specs.items():
        # name@commit: the sources of that commit (git archive into a scratch directory) instead of the working tree,
        # so that a rewritten function can be A/B-ed against what it replaced
        label, _, commit = name.partition("@")
        tmp = None
        src = os.path.join(g.CSRC, "ecmgpu.cu")
        if commit:
            tmp = tempfile.mkdtemp(prefix="ecm_variant_")
            subprocess.check_call(f"git -C {ROOT} archive {commit} ecmgenerator_b200/csrc include | tar -x -C {tmp}", shell=True)
            src = os.path.join(tmp, "ecmgenerator_b200", "csrc", "ecmgpu.cu")
        so = os.path.join(out, f"libecmgpu_{label}.so")
        cmd = [g.NVCC] + g.NVCC_FLAGS + ["-ccbin", g.GXX] + flags + ["-Xptxas", "-v", "-o", so, src, "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if tmp:
            shutil.rmtree(tmp, ignore_errors=True)
        if r.returncode != 0:
            print(r.stderr[-3000:])
            raise SystemExit(f"{name}: build failed")
        lines = r.stderr.splitlines()
        k = [i for i, l in enumerate(lines) if "Compiling entry function" in l and "k_orcaE" in l]
        used = next((l for l in lines[k[0]:k[0] + 6] if "Used" in l), "") if k else ""
        print(f"{name:24s} {' '.join(flags):40s} -> variants/{os.path.basename(so)}   k_orca:{used.split(':')[-1][:70]}")
    print("A/B:  python tools/ab_variants.py base " + " ".join(f"{n.partition('@')[0]}=variants/libecmgpu_{n.partition('@')[0]}.so" for n in specs))


if __name__ == "__main__":
    main()
