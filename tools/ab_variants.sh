#!/bin/bash
# kbench subset over every saved build under outlook_grid_vision_transformer_b200/_ab/ (OGV_LIB), two rounds.
# usage: tools/ab_variants.sh "<--only list>" "<stages>" [reps]
only=${1:-layernorm}; stages=${2:-0,1,2}; reps=${3:-20}
export PYTHONDONTWRITEBYTECODE=1
mkdir -p gpurun_out
for round in 1 2; do
  for lib in outlook_grid_vision_transformer_b200/_ab/libogvit_*.so; do
    tag=$(basename $lib .so | sed 's/libogvit_//')
    OGV_LIB=$PWD/$lib python tools/kbench.py --only "$only" --stages "$stages" --reps $reps --out gpurun_out/kb_$tag.json 2>&1 | grep "^s[0-9]" | sed "s/^/[$tag r$round] /"
  done
done
