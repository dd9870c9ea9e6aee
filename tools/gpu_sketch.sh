#!/bin/bash
# sketch-kernel round trip on the GPU box: parity tests of the fused kernel, stage times (new vs previous kernel), ncu capture
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 200 -k "sketch or fit_transform" 2>&1 | tail -15 | tee gpurun_out/pytest_sketch.log
for V in "" FDB_SKETCH_V3=1; do
  env $V timeout 600 python bench.py --config C3 --steps 5 --no-cpu-baseline --no-e2e 2>gpurun_out/sk_C3_$V.err | tee gpurun_out/sk_C3_$V.json | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('C3 [$V]', 'ms_per_step', round(d['ms_per_step'],3), {k: round(x,3) for k,x in d['stage_ms'].items()}, 'sketch frac', round(d['roofline']['sketch_kernel']['frac'],3))"
  tail -2 gpurun_out/sk_C3_$V.err
done
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"sketch_contract" -c 1 -o gpurun_out/r02_sketch_C3 -f python bench.py --config C3 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_sketch.log 2>&1
tail -3 gpurun_out/ncu_sketch.log
