#!/bin/bash
# time sweep-kernel tuning variants on one config
CFG=${1:-C3}; shift
for v in "$@"; do
  FDB_SWEEP_VARIANT=$v timeout 300 python bench.py --config $CFG --steps 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('variant $v', 'ms_per_step', round(d['ms_per_step'],2), {k: round(x,3) for k,x in d['stage_ms'].items()})"
done
