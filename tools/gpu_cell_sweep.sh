#!/usr/bin/env bash
# Neighbour-grid cell sweep on several configs (one GPU call):
#   gpurun -- 'bash tools/gpu_cell_sweep.sh <tag> "<scales>" "<configs>" "<extra variant specs>"'
# Cells are given as multiples of what build_grid (csrc/ecmgpu.cu) chooses from the crowd's local density.
set -u
TAG=${1:-cells}
OUT=gpurun_out
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
export ECM_WORKLOAD_CACHE=$PWD/workloads
SPECS="base"
for f in ${2:-1.2 1.4}; do SPECS="$SPECS s$f,AB_CELL_SCALE=$f"; done
for cfg in ${3:-c3_1m c2_50k c5_250k}; do
  for pre in 0 400; do
    echo "=== $cfg preroll $pre ($(date +%T))"
    AB_CONFIG=$cfg AB_PREROLL=$pre timeout 600 python tools/ab_variants.py $SPECS ${4:-} >"$OUT/${TAG}_cells_${cfg}_pre${pre}.jsonl" 2>>"$OUT/${TAG}_cells.err"
  done
done
echo done
