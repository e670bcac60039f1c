#!/usr/bin/env bash
# compute-sanitizer over a slice of the GPU suite (memcheck + racecheck), and an ncu capture of the device planner.
#   gpurun --timeout 1500 -- 'bash tools/gpu_hygiene.sh <tag>'
set -u
TAG=${1:-hyg}
OUT=gpurun_out
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
step() { echo "=== $1 ($(date +%T))" | tee -a "$OUT/${TAG}_session.log"; }
SLICE='tests/test_gpu_parity.py::test_lockstep_velocities_within_tolerance[jam_small] tests/test_gpu_parity.py::test_lockstep_velocities_within_tolerance[yard_small] tests/test_gpu_parity.py::test_ties_and_colocated_agents tests/test_gpu_parity.py::test_arrival_destroy_and_replan_events tests/test_gpu_parity.py::test_update_io_owned_records_match_plain_update tests/test_gpu_parity.py::test_update_io_pipeline_matches_plain_update tests/test_gpu_parity.py::test_phase_timings_are_taken_inside_the_graph_tick tests/test_gpu_parity.py::test_lockstep_velocities_within_tolerance[concave_small] tests/test_gpu_parity.py::test_nonfinite_agent_leaves_the_tick tests/test_gpu_strips.py::test_in_process_strips_match_single_gpu_bitwise tests/test_zz2_gpu_compact.py::test_compact_walk_in_the_graph_tick_with_spawns_and_destroys tests/test_zz2_gpu_spawn.py tests/test_zz4_gpu_planner.py::test_device_planner_reproduces_the_reference_polylines[c2_small] tests/test_zz4_gpu_planner.py::test_queries_that_fill_the_first_pass_scratch_are_planned_again'
step "memcheck"
timeout 900 compute-sanitizer --tool memcheck --leak-check no --error-exitcode 3 --log-file "$OUT/${TAG}_memcheck.txt" python -m pytest -m gpu -q -x $SLICE >"$OUT/${TAG}_memcheck_pytest.log" 2>&1
echo "memcheck exit $?" | tee -a "$OUT/${TAG}_session.log"; tail -2 "$OUT/${TAG}_memcheck_pytest.log"; grep -c "Invalid\|error" "$OUT/${TAG}_memcheck.txt"; tail -3 "$OUT/${TAG}_memcheck.txt"
step "racecheck"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 --log-file "$OUT/${TAG}_racecheck.txt" python -m pytest -m gpu -q -x 'tests/test_gpu_parity.py::test_lockstep_velocities_within_tolerance[jam_small]' 'tests/test_gpu_parity.py::test_lockstep_velocities_within_tolerance[concave_small]' tests/test_gpu_strips.py::test_in_process_strips_match_single_gpu_bitwise >"$OUT/${TAG}_racecheck_pytest.log" 2>&1
echo "racecheck exit $?" | tee -a "$OUT/${TAG}_session.log"; tail -2 "$OUT/${TAG}_racecheck_pytest.log"; tail -3 "$OUT/${TAG}_racecheck.txt"
if [[ " ${*:2} " != *" planner "* ]]; then step "done"; exit 0; fi
step "ncu: device planner, 200 k queries of the 1 M crowd"
export ECM_WORKLOAD_CACHE=$PWD/workloads
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_plan_paths -c 1 -o "$OUT/${TAG}_planner" -f python - >"$OUT/${TAG}_planner.log" 2>&1 <<'PY'
import time, numpy as np, bench
from ecmgenerator_b200 import gpu, scenarios as S
w, c, off, pxy = bench.build_workload("c3_1m", None)
sim = gpu.GpuSim(w, 8, float(S.DT))
n = 200_000
t = time.time()
o, p, ok = sim.plan_paths(c.pos[:n], c.goal[:n], c.radius[:n])
print("planned", n, "queries in", round(time.time() - t, 2), "s under ncu; ok", ok, "mean points", float(np.diff(o).mean()))
PY
tail -2 "$OUT/${TAG}_planner.log"
step "device planner, 1 M queries, plain timing"
timeout 300 python - >"$OUT/${TAG}_planner_timing.log" 2>&1 <<'PY'
import time, numpy as np, bench
from ecmgenerator_b200 import gpu, host, scenarios as S
w, c, off, pxy = bench.build_workload("c3_1m", None)
sim = gpu.GpuSim(w, 8, float(S.DT))
sim.plan_paths(c.pos[:1000], c.goal[:1000], c.radius[:1000])
for n in (100_000, 1_000_000):
    t = time.time(); o, p, ok = sim.plan_paths(c.pos[:n], c.goal[:n], c.radius[:n]); dt = time.time() - t
    print(f"device: {n} queries in {dt:.2f} s = {dt / n * 1e6:.2f} us per query")
t = time.time(); o2, p2, ok2 = host.plan_paths(w, c.pos[:200_000], c.goal[:200_000], c.radius[:200_000]); dt = time.time() - t
import os
print(f"host ({os.cpu_count()} cores): 200000 queries in {dt:.2f} s = {dt / 2e5 * 1e6:.2f} us per query")
PY
cat "$OUT/${TAG}_planner_timing.log" | tail -4
step "done"
