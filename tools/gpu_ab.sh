#!/usr/bin/env bash
# One short GPU call: [pytest -m gpu] + A/B of build variants (from rest and congested) + optional ncu capture.
#   gpurun --timeout 900 -- 'bash tools/gpu_ab.sh <tag> "<variant specs>" [tests] [memcheck] [ncu] [strips8]'
set -u
TAG=${1:-ab}
V=${2:-base}
OUT=gpurun_out
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
export ECM_WORKLOAD_CACHE=$PWD/workloads
step() { echo "=== $1 ($(date +%T))" | tee -a "$OUT/${TAG}_session.log"; }
if [[ " ${*:3} " == *" tests "* ]]; then
  step "pytest -m gpu"
  timeout 900 python -m pytest tests -m gpu -q -s >"$OUT/${TAG}_gpu_tests.log" 2>&1
  echo "pytest exit $?" | tee -a "$OUT/${TAG}_session.log"; tail -3 "$OUT/${TAG}_gpu_tests.log"
fi
if [[ " ${*:3} " == *" memcheck "* ]]; then
  step "compute-sanitizer memcheck over a slice of the suite"
  SLICE='tests/test_gpu_parity.py::test_lockstep_velocities_within_tolerance[jam_small] tests/test_gpu_parity.py::test_lockstep_velocities_within_tolerance[yard_small] tests/test_gpu_parity.py::test_ties_and_colocated_agents tests/test_gpu_parity.py::test_arrival_destroy_and_replan_events tests/test_gpu_parity.py::test_update_io_owned_records_match_plain_update tests/test_gpu_parity.py::test_nonfinite_agent_leaves_the_tick tests/test_gpu_strips.py::test_in_process_strips_match_single_gpu_bitwise tests/test_gpu_strips.py::test_automatic_rebalancing_keeps_results_bitwise tests/test_zz2_gpu_compact.py::test_compact_walk_in_the_graph_tick_with_spawns_and_destroys tests/test_zz2_gpu_spawn.py tests/test_zz4_gpu_planner.py::test_device_planner_reproduces_the_reference_polylines[c2_small]'
  timeout 900 compute-sanitizer --tool memcheck --leak-check no --error-exitcode 3 --log-file "$OUT/${TAG}_memcheck.txt" python -m pytest -m gpu -q $SLICE >"$OUT/${TAG}_memcheck_pytest.log" 2>&1
  echo "memcheck exit $?" | tee -a "$OUT/${TAG}_session.log"; tail -2 "$OUT/${TAG}_memcheck_pytest.log"; tail -2 "$OUT/${TAG}_memcheck.txt"
fi
step "A/B from rest"
timeout 600 python tools/ab_variants.py $V >"$OUT/${TAG}_ab_rest.jsonl" 2>"$OUT/${TAG}_ab_rest.err"
step "A/B congested (400 ticks of pre-roll)"
AB_PREROLL=400 timeout 600 python tools/ab_variants.py $V >"$OUT/${TAG}_ab_congested.jsonl" 2>"$OUT/${TAG}_ab_congested.err"
if [[ " ${*:3} " == *" ncu "* ]]; then
  step "ncu launch list"
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file "$OUT/${TAG}_launches.csv" \
      python bench.py --steps 2 --warmup 3 --no-cpu --steady-tick 0 >"$OUT/${TAG}_ncu_launch_bench.log" 2>&1
  step "ncu --set full"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_orca|k_attract|k_scatter|k_bin_count" --launch-skip 12 -c 4 \
      -o "$OUT/${TAG}_full" -f python bench.py --steps 3 --warmup 5 --no-cpu --steady-tick 0 >"$OUT/${TAG}_ncu_full_bench.log" 2>&1
fi
if [[ " ${*:3} " == *" strips8 "* ]]; then
  step "ncu launch list of 8 in-process strips (per-rank problem size of an 8-GPU run)"
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT/${TAG}_strips8_launches.csv" \
      python tools/strip_profile.py --strips 8 --ticks 2 >"$OUT/${TAG}_strips8.log" 2>&1
fi
step "done"
