#!/usr/bin/env bash
# One gpurun call that measures everything prepared on the CPU (run from the repo root on a B200 box):
#
#   gpurun --timeout 3600 -- 'bash tools/gpu_session.sh r02a'      (about 40 GPU-minutes; every step has its own timeout)
#
# Steps (each under its own timeout, all output under gpurun_out/<tag>_*):
#   1. pytest -m gpu                      the parity suite incl. the KD-tree mode and the full-size windows
#   2. bench.py                           the headline line (1 GPU, c3_1m)
#   3. tools/ab_variants.py               base vs the prepared build variants, one process, same box
#   4. the same on the congested crowd    (AB_PREROLL=400)
#   5. cell-size sweep for the pruning variants (smaller cells pay only with pruning)
#   6. parity statistics of the orca_fast variant (NOT bit-exact by design)
#   7. bench.py --neighbors kdtree        cost of the parity mode
#   7b. bench.py --planner device         set-up with ecmgpu_plan_paths instead of the host planner
#   7c. bench.py at 5 k / 50 k / 250 k / 4 M agents on one GPU
#   7d. per-kernel cost of a strip tick (8 in-process strips) with and without the compact walk
#   8. ncu launch list + one --set full capture of k_orca / k_attract of the default build
# Nothing here changes GPU clocks.  Numbers printed under ncu are never bench values.
set -u
TAG=${1:-session}
OUT=gpurun_out
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
export ECM_WORKLOAD_CACHE=/tmp/ecm_workloads   # the 1 M crowd's routes are planned once per session, not once per step
step() { echo "=== $1 ($(date +%T))" | tee -a "$OUT/${TAG}_session.log"; }

step "build"
python -c "import __graft_entry__ as g; g.build()" >>"$OUT/${TAG}_session.log" 2>&1
python tools/build_variants.py >"$OUT/${TAG}_variants_build.log" 2>&1

step "pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q -s -x >"$OUT/${TAG}_gpu_tests.log" 2>&1
echo "pytest exit $?" | tee -a "$OUT/${TAG}_session.log"

step "bench"
timeout 600 python bench.py >"$OUT/${TAG}_bench.json" 2>"$OUT/${TAG}_bench.err"

# name[=library][,ENV=VALUE]: the split tick is a run-time switch of the default library (and of any variant)
V="base attract_bbox4=variants/libecmgpu_attract_bbox4.so split,ECMGPU_SPLIT=1 split_fast=variants/libecmgpu_orca_fast.so,ECMGPU_SPLIT=1 split_twopass=variants/libecmgpu_knn_twopass.so,ECMGPU_SPLIT=1 knn_flat=variants/libecmgpu_knn_flat.so split_flat=variants/libecmgpu_knn_flat.so,ECMGPU_SPLIT=1 knn_flat_prune=variants/libecmgpu_knn_flat_prune.so knn_prune=variants/libecmgpu_knn_prune.so knn_twopass=variants/libecmgpu_knn_twopass.so knn_twopass_prune=variants/libecmgpu_knn_twopass_prune.so orca_fast=variants/libecmgpu_orca_fast.so"
step "A/B from rest"
timeout 900 python tools/ab_variants.py $V >"$OUT/${TAG}_ab_rest.jsonl" 2>"$OUT/${TAG}_ab_rest.err"
step "A/B congested (400 ticks of pre-roll)"
AB_PREROLL=400 timeout 1200 python tools/ab_variants.py $V >"$OUT/${TAG}_ab_congested.jsonl" 2>"$OUT/${TAG}_ab_congested.err"

step "cell sweep with pruning"
: >"$OUT/${TAG}_ab_cells.jsonl"
for cell in 2.0 2.4 2.8; do
  AB_CELL=$cell timeout 600 python tools/ab_variants.py base knn_prune=variants/libecmgpu_knn_prune.so knn_twopass_prune=variants/libecmgpu_knn_twopass_prune.so \
      >>"$OUT/${TAG}_ab_cells.jsonl" 2>>"$OUT/${TAG}_ab_cells.err"
done

step "orca_fast parity statistics"
ECMGPU_LIB=$PWD/variants/libecmgpu_orca_fast.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s \
    -k "lockstep or free_running or rms_600" >"$OUT/${TAG}_orca_fast_parity.log" 2>&1

step "bench --neighbors kdtree"
timeout 600 python bench.py --neighbors kdtree --no-cpu --steady-tick 0 --steps 20 >"$OUT/${TAG}_bench_kdtree.json" 2>"$OUT/${TAG}_bench_kdtree.err"

step "bench --planner device (set-up time on stderr)"
ECM_WORKLOAD_CACHE= timeout 900 python bench.py --planner device --no-cpu --steady-tick 0 --steps 20 >"$OUT/${TAG}_bench_devplan.json" 2>"$OUT/${TAG}_bench_devplan.err"

step "bench at the other BASELINE sizes (SURVEY.md 8d: 5 k, 50 k, 250 k on one GPU; 4 M with device-planned routes)"
for cfg in c1_5k c2_50k c5_250k; do
  ECM_WORKLOAD_CACHE= timeout 600 python bench.py --config $cfg --no-cpu --steady-tick 0 >"$OUT/${TAG}_bench_${cfg}.json" 2>"$OUT/${TAG}_bench_${cfg}.err"
done
ECM_WORKLOAD_CACHE= timeout 1200 python bench.py --config c4_4m --planner device --no-cpu --steady-tick 0 --steps 50 \
    >"$OUT/${TAG}_bench_c4_4m.json" 2>"$OUT/${TAG}_bench_c4_4m.err"

step "strip-scale launch lists on one GPU: 8 in-process strips, all-slots walk vs ECMGPU_COMPACT=1"
for c in 0 1; do
  ECMGPU_COMPACT=$c timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
      --log-file "$OUT/${TAG}_strips8_compact${c}_launches.csv" python tools/strip_profile.py --strips 8 --ticks 2 \
      >"$OUT/${TAG}_strips8_compact${c}.log" 2>&1
done

step "ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/${TAG}_launches.csv" \
    python bench.py --steps 2 --warmup 3 --no-cpu --steady-tick 0 >"$OUT/${TAG}_ncu_launch_bench.log" 2>&1
step "ncu --set full (k_orca, k_attract)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_orca|k_attract|k_scatter" --launch-skip 15 -c 6 \
    -o "$OUT/${TAG}_full" -f python bench.py --steps 4 --warmup 5 --no-cpu --steady-tick 0 >"$OUT/${TAG}_ncu_full_bench.log" 2>&1
step "done"
