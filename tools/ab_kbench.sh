#!/bin/bash
# A/B of a kernel change on ONE box: the same kbench subset with the saved baseline build
# (outlook_grid_vision_transformer_b200/_ab/libogvit_base.so, loaded through OGV_LIB) and with the current build.
# usage: tools/ab_kbench.sh "<--only list>" "<stages>" [reps]
only=${1:-dwconv,gemm}; stages=${2:-0,1}; reps=${3:-20}
export PYTHONDONTWRITEBYTECODE=1
mkdir -p gpurun_out
BASE=$PWD/outlook_grid_vision_transformer_b200/_ab/libogvit_base.so
for round in 1 2; do
  echo "== base (round $round)"; OGV_LIB=$BASE python tools/kbench.py --only "$only" --stages "$stages" --reps $reps --out gpurun_out/kb_base.json 2>&1 | grep -v "^$"
  echo "== new  (round $round)"; python tools/kbench.py --only "$only" --stages "$stages" --reps $reps --out gpurun_out/kb_new.json 2>&1 | grep -v "^$"
done
