#!/usr/bin/env bash
# Multi-GPU companion of tools/gpu_session.sh (run from the repo root on a box with N GPUs):
#
#   gpurun --gpus 4 --timeout 1800 -- 'bash tools/gpu_session_multi.sh r02m 4'
#
#   1. the strips tests (in-process, multi-process over NCCL and the peer transport, compact walk)
#   2. bench.py on N GPUs, default strips and ECMGPU_COMPACT=1 (pack / count / scatter / attract / orca / grid scale with the
#      rank's share), 1 M agents strong scaling; then the 4 M map with device-planned routes
# Output under gpurun_out/<tag>_*.  Multi-rank commands are never wrapped in ncu.
set -u
TAG=${1:-multi}
N=${2:-4}
OUT=gpurun_out
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
step() { echo "=== $1 ($(date +%T))" | tee -a "$OUT/${TAG}_session.log"; }
run_bench() {  # name, extra env, bench args...
  local name=$1 envs=$2; shift 2
  env $envs timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus "$N" "$@" >"$OUT/${TAG}_${name}.json" 2>"$OUT/${TAG}_${name}.err"
}
step "build"
python -c "import __graft_entry__ as g; g.build()" >>"$OUT/${TAG}_session.log" 2>&1
step "strips tests"
timeout 900 python -m pytest tests/test_gpu_strips.py tests/test_zz2_gpu_split.py -m gpu -q -s >"$OUT/${TAG}_strips_tests.log" 2>&1
echo "pytest exit $?" | tee -a "$OUT/${TAG}_session.log"
step "strips tests with ECMGPU_COMPACT=1 (every transport with the compact walk)"
ECMGPU_COMPACT=1 timeout 900 python -m pytest tests/test_gpu_strips.py -m gpu -q -s >"$OUT/${TAG}_strips_tests_compact.log" 2>&1
echo "pytest exit $?" | tee -a "$OUT/${TAG}_session.log"
step "bench c3_1m x$N, default strips"
run_bench bench_n${N} "ECMGPU_COMPACT=0" --steady-tick 0
step "bench c3_1m x$N, compact walk"
run_bench bench_n${N}_compact "ECMGPU_COMPACT=1" --steady-tick 0
step "bench c3_1m x$N, compact walk + split tick"
run_bench bench_n${N}_compact_split "ECMGPU_COMPACT=1 ECMGPU_SPLIT=1" --steady-tick 0
step "bench c4_4m x$N (routes planned on the GPUs), default and compact"
run_bench bench_c4_n${N} "ECMGPU_COMPACT=0" --config c4_4m --planner device --steady-tick 0 --steps 50
run_bench bench_c4_n${N}_compact "ECMGPU_COMPACT=1" --config c4_4m --planner device --steady-tick 0 --steps 50
step "done"
