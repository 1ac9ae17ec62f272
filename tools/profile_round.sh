#!/bin/bash
# Evidence for one measurement round, on the GPU box:
#   1. ncu launch list of the bench command (gpu__time_duration per launch; cold-cache, serialised -> SHARES)
#   2. one `ncu --set full` capture of the dominant kernel's launches inside the same command
# usage: tools/profile_round.sh <tag> <kernel regex>      e.g.  tools/profile_round.sh r01c dwconv_bwd
tag=${1:-r01}; top=${2:-dwconv_bwd}
export PYTHONDONTWRITEBYTECODE=1
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --no-graph --no-profile --no-cpu-baseline"
$BENCH > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain bench run failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
tail -n 1 gpurun_out/${tag}_plain.log | cut -c1-160
ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/${tag}_launches.csv $BENCH > gpurun_out/${tag}_ncu_list.log 2>&1
echo "launch list rc=$? rows=$(wc -l < gpurun_out/${tag}_launches.csv)"
ncu --set full --clock-control none --import-source on -k regex:"$top" -c 8 -f -o gpurun_out/${tag}_top $BENCH > gpurun_out/${tag}_ncu_full.log 2>&1
echo "full capture rc=$?"
ncu -i gpurun_out/${tag}_top.ncu-rep --page raw --csv > gpurun_out/${tag}_top.raw.csv 2>/dev/null
ls -la gpurun_out/${tag}_top.ncu-rep
# keep the merged directory under the 64 MiB copy-back limit
if [ $(stat -c %s gpurun_out/${tag}_top.ncu-rep) -gt 40000000 ]; then rm -f gpurun_out/${tag}_top.ncu-rep; echo "(report too large to copy back; raw CSV kept)"; fi
du -sh gpurun_out
