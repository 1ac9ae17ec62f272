#!/bin/bash
# One GPU-box visit: parity suite, smoke, bench, then an ncu launch list of the same bench command.
# usage: tools/gpu_check.sh [tests|bench|ncu ...]   (default: all three)
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
what="${*:-tests bench ncu}"
rc_all=0
if [[ "$what" == *tests* ]]; then
  rm -f gpurun_out/suite.log
  bash tools/run_gpu_suite.sh tests/test_gpu_kernels.py tests/test_gpu_gemm.py tests/test_gpu_parity_golden.py tests/test_gpu_block_oracle.py $OGV_EXTRA_TESTS || rc_all=1
  timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1 || rc_all=1
  tail -n 4 gpurun_out/smoke.log
fi
if [[ "$what" == *bench* ]]; then
  timeout 900 python bench.py --steps ${OGV_BENCH_STEPS:-10} --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err || rc_all=1
  tail -n 45 gpurun_out/bench.err
  tail -n 1 gpurun_out/bench.log
fi
if [[ "$what" == *ncu* ]]; then
  BENCH="python bench.py --steps 1 --warmup 3 --no-graph --no-profile --no-cpu-baseline"
  timeout 600 $BENCH > gpurun_out/ncu_plain.log 2>&1 &&
  timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'ogv|gemm|ln_|bn_|dw|outlook|grid|colsum|colstats|se_|kernel' -c 20000 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu.log 2>&1 || rc_all=1
  tail -n 3 gpurun_out/ncu.log
fi
exit $rc_all
