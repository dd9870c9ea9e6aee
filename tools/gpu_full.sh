#!/bin/bash
# full GPU round trip: parity suite, then the default bench line (C3 + parity + c5)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
timeout 1500 python bench.py --steps ${STEPS:-5} ${BENCH_ARGS} 2>gpurun_out/bench_default.err | tee gpurun_out/bench_default.json | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('ms_per_step', round(d['ms_per_step'],3), {k: round(x,3) for k,x in d['stage_ms'].items()}, 'sweep frac', round(d['roofline']['frac'],3), 'sketch frac', round(d['roofline']['sketch_kernel']['frac'],3))
print('e2e', d['e2e']); print('parity', d.get('parity')); print('cpu', d.get('cpu_baseline',{}).get('value')); print('c5', d.get('c5'))"
tail -3 gpurun_out/bench_default.err
