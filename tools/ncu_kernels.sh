#!/bin/bash
# One `ncu --set full` capture (the timed launch only) per kernel of interest, driven through
# tools/kbench.py; exports the raw metric page as CSV next to each report.
#   usage: tools/ncu_kernels.sh <tag> "<stage>:<kbench --only substring>:<kernel regex>" ...
export PYTHONDONTWRITEBYTECODE=1
tag=$1; shift
mkdir -p gpurun_out
for spec in "$@"; do
  IFS=: read -r stage only regex <<< "$spec"
  name=$(echo "${tag}_s${stage}_${only}" | tr -c 'A-Za-z0-9_\n' '_')
  CMD=(python tools/kbench.py --stages "$stage" --only "$only" --reps 1 --out gpurun_out/kb_tmp.json)
  "${CMD[@]}" > gpurun_out/plain_$name.log 2>&1 || { echo "plain run failed for $spec"; continue; }
  ncu --set full --clock-control none --import-source on -k regex:"$regex" -s 3 -c 1 -f -o gpurun_out/$name "${CMD[@]}" > gpurun_out/ncu_$name.log 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
  ls -la gpurun_out/$name.ncu-rep
done
