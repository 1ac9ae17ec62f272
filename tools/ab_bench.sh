#!/bin/bash
# A/B of environment switches on ONE box: tools/ab_bench.sh "NAME=VAL ..." "NAME=VAL ..." ...   ("" = defaults)
# prints ms/step (resident) and e2e ms/step of the headline workload for every variant, twice round-robin.
FLAGS="--no-cpu-baseline --no-profile --no-eager-gpu --no-extra ${AB_FLAGS:-}"
for rep in 1 2; do
  for v in "$@"; do
    out=$(env $v python bench.py $FLAGS 2>/dev/null | tail -1)
    python - "$v" "$out" <<'PY'
import json, sys
d = json.loads(sys.argv[2])
print(f"{sys.argv[1] or 'default':45s} {d['ms_per_step']:8.3f} ms  e2e {d['e2e']['ms_per_step']:8.3f} ms  launches {d.get('launches_per_step')}")
PY
  done
done
