#!/usr/bin/env bash
# GPU tests + the bench line of every BASELINE config on one B200.
#   gpurun --timeout 2400 -- 'bash tools/gpu_bench_matrix.sh <tag> [tests] [c4]'
set -u
TAG=${1:-mx}
OUT=gpurun_out
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
export ECM_WORKLOAD_CACHE=$PWD/workloads
step() { echo "=== $1 ($(date +%T))" | tee -a "$OUT/${TAG}_session.log"; }
if [[ " ${*:2} " == *" tests "* ]]; then
  step "pytest -m gpu"
  timeout 1200 python -m pytest tests -m gpu -q -x -s >"$OUT/${TAG}_gpu_tests.log" 2>&1
  echo "pytest exit $?" | tee -a "$OUT/${TAG}_session.log"; tail -3 "$OUT/${TAG}_gpu_tests.log"
fi
step "bench c3_1m (default line)"
timeout 900 python bench.py >"$OUT/${TAG}_bench_c3_1m.json" 2>"$OUT/${TAG}_bench_c3_1m.err"
for cfg in c1_5k c2_50k c5_250k; do
  step "bench $cfg"
  timeout 600 python bench.py --config $cfg >"$OUT/${TAG}_bench_${cfg}.json" 2>"$OUT/${TAG}_bench_${cfg}.err"
done
if [[ " ${*:2} " == *" c4 "* ]]; then
  step "bench c4_4m on one GPU (routes planned on the device)"
  ECM_WORKLOAD_CACHE= timeout 1200 python bench.py --config c4_4m --planner device --no-cpu --steps 50 >"$OUT/${TAG}_bench_c4_4m.json" 2>"$OUT/${TAG}_bench_c4_4m.err"
fi
step "done"
