#!/bin/bash
# Run the -m gpu test files one process each (a CUDA fault in one file cannot poison the others),
# every file under its own timeout; logs land in gpurun_out/.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
status=0
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for f in "$@"; do
  name=$(basename "$f" .py)
  echo "=== $f" | tee -a gpurun_out/suite.log
  timeout ${OGV_TEST_TIMEOUT:-600} python -m pytest "$f" -q -m gpu -p no:cacheprovider --tb=short --maxfail=40 --timeout 300 ${OGV_PYTEST_ARGS} > "gpurun_out/$name.log" 2>&1
  rc=$?
  tail -n 25 "gpurun_out/$name.log" | tee -a gpurun_out/suite.log
  echo "rc=$rc" | tee -a gpurun_out/suite.log
  [ $rc -ne 0 ] && status=1
done
exit $status
