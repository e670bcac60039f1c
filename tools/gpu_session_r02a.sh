#!/usr/bin/env bash
# Round-2 first session (one B200, ~15 GPU-minutes): the A/B of the build variants prepared in round 1, the parity
# statistics of orca_fast, bench lines at the small configs, the two parity/setup modes, strip-scale launch lists.
set -u
TAG=${1:-r02a}
OUT=gpurun_out
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
export ECM_WORKLOAD_CACHE=$PWD/workloads
step() { echo "=== $1 ($(date +%T))" | tee -a "$OUT/${TAG}_session.log"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv >>"$OUT/${TAG}_session.log" 2>&1
V="base attract_bbox4=variants/libecmgpu_attract_bbox4.so split,ECMGPU_SPLIT=1 split_fast=variants/libecmgpu_orca_fast.so,ECMGPU_SPLIT=1 split_twopass=variants/libecmgpu_knn_twopass.so,ECMGPU_SPLIT=1 knn_flat=variants/libecmgpu_knn_flat.so knn_flat_prune=variants/libecmgpu_knn_flat_prune.so knn_prune=variants/libecmgpu_knn_prune.so knn_twopass=variants/libecmgpu_knn_twopass.so knn_twopass_prune=variants/libecmgpu_knn_twopass_prune.so orca_fast=variants/libecmgpu_orca_fast.so base2"
step "A/B from rest"
timeout 600 python tools/ab_variants.py $V >"$OUT/${TAG}_ab_rest.jsonl" 2>"$OUT/${TAG}_ab_rest.err"
step "A/B congested (400 ticks of pre-roll)"
AB_PREROLL=400 timeout 600 python tools/ab_variants.py $V >"$OUT/${TAG}_ab_congested.jsonl" 2>"$OUT/${TAG}_ab_congested.err"
step "cell 2.4 with pruning"
AB_CELL=2.4 timeout 300 python tools/ab_variants.py base knn_prune=variants/libecmgpu_knn_prune.so knn_twopass_prune=variants/libecmgpu_knn_twopass_prune.so \
      >"$OUT/${TAG}_ab_cells.jsonl" 2>"$OUT/${TAG}_ab_cells.err"
step "orca_fast parity statistics"
ECMGPU_LIB=$PWD/variants/libecmgpu_orca_fast.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s \
    -k "lockstep or free_running or rms_600" >"$OUT/${TAG}_orca_fast_parity.log" 2>&1
step "bench --neighbors kdtree"
timeout 300 python bench.py --neighbors kdtree --no-cpu --steady-tick 0 --steps 20 >"$OUT/${TAG}_bench_kdtree.json" 2>"$OUT/${TAG}_bench_kdtree.err"
step "bench small configs"
for cfg in c1_5k c2_50k c5_250k; do
  ECM_WORKLOAD_CACHE= timeout 300 python bench.py --config $cfg --no-cpu --steady-tick 0 >"$OUT/${TAG}_bench_${cfg}.json" 2>"$OUT/${TAG}_bench_${cfg}.err"
done
step "bench --planner device (set-up time on stderr)"
ECM_WORKLOAD_CACHE= timeout 400 python bench.py --planner device --no-cpu --steady-tick 0 --steps 20 >"$OUT/${TAG}_bench_devplan.json" 2>"$OUT/${TAG}_bench_devplan.err"
step "strip-scale launch lists: 8 in-process strips, all-slots walk vs ECMGPU_COMPACT=1"
for c in 0 1; do
  ECMGPU_COMPACT=$c timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
      --log-file "$OUT/${TAG}_strips8_compact${c}_launches.csv" python tools/strip_profile.py --strips 8 --ticks 2 \
      >"$OUT/${TAG}_strips8_compact${c}.log" 2>&1
done
step "done"
