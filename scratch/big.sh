#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 200 2>&1 | tail -4
for CFG in C4 C5; do
timeout 900 python bench.py --config $CFG --steps 2 --no-cpu-baseline > gpurun_out/bench_$CFG.json 2> gpurun_out/bench_$CFG.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_$CFG.json').read().strip().splitlines()[-1])
    print('$CFG', 'ms/step', round(d['ms_per_step'],2), 'value', round(d['value']), {k: round(v,3) for k,v in d['stage_ms'].items()}, 'roof', round(d['roofline']['frac'],3), 'sketch', round(d['roofline']['sketch_kernel']['frac'],3), 'e2e ms', round(d['e2e']['ms_per_step'],1), 'deg', round(d['config']['mean_degree'],2), 'nnz', d['config']['nnz'])
except Exception as ex:
    print('$CFG failed', ex); print(open('gpurun_out/bench_$CFG.err').read()[-1500:])
PY
done
