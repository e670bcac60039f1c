#!/usr/bin/env bash
# Device planner: queries in flight (launch bounds variants from variants/libecmgpu_pl*.so) and first-pass push capacity.
#   gpurun -- 'bash tools/gpu_planner_sweep.sh <tag>'
set -u
TAG=${1:-plan}
OUT=gpurun_out
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
export ECM_WORKLOAD_CACHE=$PWD/workloads
N=${2:-300000}
for lib in "" variants/libecmgpu_pl8.so variants/libecmgpu_pl10.so variants/libecmgpu_pl12.so; do
  echo "=== lib ${lib:-default}" | tee -a "$OUT/${TAG}_planner_sweep.jsonl"
  ECMGPU_LIB=${lib:+$PWD/$lib} timeout 300 python tools/planner_probe.py $N 0 98304 >>"$OUT/${TAG}_planner_sweep.jsonl" 2>>"$OUT/${TAG}_planner_sweep.err"
done
for push in 2048 4096 16384; do
  echo "=== push $push" | tee -a "$OUT/${TAG}_planner_sweep.jsonl"
  ECMGPU_PLAN_PUSH=$push timeout 300 python tools/planner_probe.py $N 0 >>"$OUT/${TAG}_planner_sweep.jsonl" 2>>"$OUT/${TAG}_planner_sweep.err"
done
echo "=== python -m pytest planner tests"
timeout 600 python -m pytest tests/test_zz4_gpu_planner.py -m gpu -q 2>&1 | tail -3
echo done
