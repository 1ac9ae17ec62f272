#!/bin/bash
# Evidence for one measurement round, on the GPU box (bench command = the step run eagerly, no graph):
#   1. ncu launch list of the bench command (gpu__time_duration per launch; cold-cache, serialised -> SHARES)
#   2. DRAM bytes + duration of EVERY launch of the dominant kernel (metrics-only pass) -> traffic per launch
#   3. one `ncu --set full` capture of a few launches of the dominant kernel (source-level stalls)
# usage: [SKIP_LIST=1] tools/profile_round.sh <tag> <kernel regex> [launches for the full capture]
#   <kernel regex> matches the FUNCTION name only (ncu -k ignores template arguments): e.g. gemm_tc_kernel
tag=${1:-r01}; top=${2:-gemm_tc_kernel}; nfull=${3:-6}
export PYTHONDONTWRITEBYTECODE=1
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --no-graph --no-profile --no-cpu-baseline --no-eager-gpu --no-extra"
$BENCH > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain bench run failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
tail -n 1 gpurun_out/${tag}_plain.log | cut -c1-160
if [ -z "$SKIP_LIST" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file gpurun_out/${tag}_launches.csv $BENCH > gpurun_out/${tag}_ncu_list.log 2>&1
  echo "launch list rc=$? rows=$(wc -l < gpurun_out/${tag}_launches.csv)"
fi
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"$top" -c 40000 --csv --log-file gpurun_out/${tag}_top_traffic.csv $BENCH > gpurun_out/${tag}_ncu_traffic.log 2>&1
echo "traffic pass rc=$? rows=$(wc -l < gpurun_out/${tag}_top_traffic.csv)"
ncu --set full --clock-control none --import-source on -k regex:"$top" -s 40 -c $nfull -f -o gpurun_out/${tag}_top $BENCH > gpurun_out/${tag}_ncu_full.log 2>&1
echo "full capture rc=$?"
ncu -i gpurun_out/${tag}_top.ncu-rep --page raw --csv > gpurun_out/${tag}_top.raw.csv 2>/dev/null
ls -la gpurun_out/${tag}_top.ncu-rep
if [ -f gpurun_out/${tag}_top.ncu-rep ] && [ $(stat -c %s gpurun_out/${tag}_top.ncu-rep) -gt 30000000 ]; then rm -f gpurun_out/${tag}_top.ncu-rep; echo "(report too large to copy back; raw CSV kept)"; fi
du -sh gpurun_out
