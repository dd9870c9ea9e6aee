#!/bin/bash
# quick GPU check: parity tests, then stage times on the given configs (VARIANTS="0 4" compares sweep kernels)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 200 2>&1 | tail -4
for CFG in "$@"; do
  for V in ${VARIANTS:-0}; do
  FDB_SWEEP_VARIANT=$V timeout 600 python bench.py --config $CFG --steps 3 --no-cpu-baseline --no-e2e 2>gpurun_out/quick_$CFG.err | tee gpurun_out/quick_${CFG}_v$V.json | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('$CFG variant $V', 'ms_per_step', round(d['ms_per_step'],3), {k: round(x,3) for k,x in d['stage_ms'].items()}, 'roof', round(d['roofline']['frac'],3), d['roofline'].get('kernel'))"
  tail -2 gpurun_out/quick_$CFG.err
  done
done
