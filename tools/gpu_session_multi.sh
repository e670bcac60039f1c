#!/usr/bin/env bash
# Multi-GPU session (run from the repo root on a box with N GPUs):
#   gpurun --gpus 2 --timeout 1800 -- 'bash tools/gpu_session_multi.sh r02m 2 [tests] [c4]'
#   1. the strips tests (in-process, multi-process over NCCL and the peer transport)
#   2. bench.py on N GPUs: 1 M agents strong scaling with the bitwise-vs-1-GPU parity field; optionally the 4 M map
# Output under gpurun_out/<tag>_*.  Multi-rank commands are never wrapped in ncu.
set -u
TAG=${1:-multi}
N=${2:-2}
OUT=gpurun_out
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
export ECM_WORKLOAD_CACHE=$PWD/workloads
step() { echo "=== $1 ($(date +%T))" | tee -a "$OUT/${TAG}_session.log"; }
run_bench() {  # name, extra env, bench args...
  local name=$1 envs=$2; shift 2
  env $envs timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus "$N" "$@" >"$OUT/${TAG}_${name}.json" 2>"$OUT/${TAG}_${name}.err"
  echo "bench $name exit $?" | tee -a "$OUT/${TAG}_session.log"
}
nvidia-smi -L >>"$OUT/${TAG}_session.log" 2>&1
if [[ " ${*:3} " == *" tests "* ]]; then
  step "strips tests"
  timeout 1200 python -m pytest tests -m gpu -q -s >"$OUT/${TAG}_strips_tests.log" 2>&1
  echo "pytest exit $?" | tee -a "$OUT/${TAG}_session.log"; tail -3 "$OUT/${TAG}_strips_tests.log"
fi
if [[ " ${*:3} " != *" noc3 "* ]]; then
  step "bench c3_1m x$N"
  run_bench bench_n${N} "X=1"
fi
if [[ " ${*:3} " == *" c4 "* ]]; then
  step "bench c4_4m x$N (routes planned on the GPUs)"
  ECM_WORKLOAD_CACHE= run_bench bench_c4_n${N} "X=1" --config c4_4m --planner device --steps 50
fi
step "done"
