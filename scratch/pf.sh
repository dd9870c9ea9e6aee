#!/bin/bash
for pf in 0 444 222 888; do
  FDB_SWEEP_PREFETCH=$pf timeout 300 python bench.py --config C3 --steps 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('prefetch $pf', 'ms_per_step', round(d['ms_per_step'],2), {k: round(x,3) for k,x in d['stage_ms'].items()})"
done
