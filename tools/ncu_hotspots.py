"""Source lines of one kernel ordered by stall samples, with their share of warp instructions and lanes per instruction.

  python tools/ncu_hotspots.py report.ncu-rep k_orca [lines]
"""
import collections
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, agg, files = "?", None, [], collections.Counter()
for r in rows:
    if len(r) >= 2 and r[0].strip() == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) > 10:
        try:
            ln = int(r[0])
        except ValueError:
            continue
        d = dict(zip(hdr[4:], r[4:]))
        if not d["# Samples"].isdigit():
            continue
        agg.append((cur, ln, r[1].strip(), int(d["# Samples"]), int(d["Instructions Executed"]), int(d["Thread Instructions Executed"])))
        files[cur] += int(d["# Samples"])
ts, ti, tt = sum(a[3] for a in agg), sum(a[4] for a in agg), sum(a[5] for a in agg)
print(f"total samples {ts} warp-instr {ti} thread-instr {tt} simt-eff {tt / max(1, ti) / 32:.2f}")
print("samples by file:", dict(files))
for f, ln, src, s, i, t in sorted(agg, key=lambda a: -a[3])[:top]:
    print(f"{f:12s}:{ln:4d} smp {s / ts:5.1%} inst {i / ti:5.1%} thr/inst {t / max(1, i):4.1f} | {src[:90]}")
