#!/bin/bash
# usage: scale.sh N CFG [CFG...]
N=$1; shift
mkdir -p gpurun_out
for CFG in "$@"; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus $N --steps 3 --warmup 3 --config $CFG > gpurun_out/bench_${CFG}_g$N.json 2> gpurun_out/bench_${CFG}_g$N.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${CFG}_g$N.json').read().strip().splitlines()[-1])
    print('N=$N $CFG', 'ms/step', round(d['ms_per_step'],2), 'value', round(d['value']), {k: round(v,3) for k,v in d['stage_ms'].items()}, d['config']['multi_gpu'][:90], 'e2e ms', round(d['e2e']['ms_per_step'],1), 'obj', d['final_objective'])
except Exception as ex:
    print('failed', ex); print(open('gpurun_out/bench_${CFG}_g$N.err').read()[-2000:])
PY
done
