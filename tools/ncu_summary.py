"""Turns an `ncu --set full` report into the two files bench.py and the judge read:

  python tools/ncu_summary.py gpurun_out/prof_v8.ncu-rep profiles/r01_v8 "source description"

  <prefix>_ncu_full_raw.csv   `ncu -i ... --page raw --csv` (every metric of every captured launch)
  <prefix>_traffic.json       DRAM bytes and duration per kernel (mean over its captured launches) - bench.py's
                              roofline.traffic reads the newest profiles/*_traffic.json
and prints the scheduler / stall metrics DESIGN.md quotes.
"""
import collections
import csv
import io
import json
import subprocess
import sys


def main():
    rep, prefix = sys.argv[1], sys.argv[2]
    src = sys.argv[3] if len(sys.argv) > 3 else rep
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    open(prefix + "_ncu_full_raw.csv", "w").write(raw)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name):
        try:
            return float(r[col[name]].replace(",", ""))
        except (KeyError, ValueError):
            return float("nan")

    def scale(name, to):
        u = units[col[name]]
        k = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
        return k.get(u, 1.0) / (1e6 if to == "MB" else 1.0)

    per = collections.OrderedDict()
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0]
        per.setdefault(name, []).append(r)
    out = {"source": src, "kernels": {}}
    watch = ["smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
             "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
             "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
             "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
             "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
             "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
             "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
             "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
             "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
             "launch__registers_per_thread", "launch__occupancy_limit_registers"]
    for name, rs in per.items():
        n = len(rs)
        k = {"launches": n,
             "dram_read_mbytes": sum(val(r, "dram__bytes_read.sum") for r in rs) / n * scale("dram__bytes_read.sum", "MB"),
             "dram_write_mbytes": sum(val(r, "dram__bytes_write.sum") for r in rs) / n * scale("dram__bytes_write.sum", "MB"),
             "duration_us": sum(val(r, "gpu__time_duration.sum") for r in rs) / n * scale("gpu__time_duration.sum", "us")}
        out["kernels"][name] = k
        print(name, json.dumps({a: round(b, 3) if isinstance(b, float) else b for a, b in k.items()}))
        for w in watch:
            if w in col:
                print(f"    {w:82s} {sum(val(r, w) for r in rs) / n:12.3f} {units[col[w]]}")
    json.dump(out, open(prefix + "_traffic.json", "w"), indent=1)


if __name__ == "__main__":
    main()
