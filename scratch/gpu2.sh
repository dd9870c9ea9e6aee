#!/bin/bash
N=${1:-2}
CFG=${2:-C3}
mkdir -p gpurun_out
for MODE in peer nccl; do FDB_TILED_MODE=$MODE timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/check_tiled.py 20000 2>&1 | grep -E "tiled|Error|error|warn" | tail -5; done
FDB_TILED_TORCH=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 tools/check_tiled.py 300000 2>&1 | grep -E "tiled|Error|error" | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $N --steps 3 --warmup 3 --config $CFG > gpurun_out/bench_${CFG}_g$N.json 2> gpurun_out/bench_${CFG}_g$N.err; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${CFG}_g$N.json').read().strip().splitlines()[-1])
print('N=$N', 'ms/step', round(d['ms_per_step'],2), 'value', round(d['value']), {k: round(v,3) for k,v in d['stage_ms'].items()}, d['config']['multi_gpu'], 'e2e ms', round(d['e2e']['ms_per_step'],1))
PY
tail -3 gpurun_out/bench_${CFG}_g$N.err
