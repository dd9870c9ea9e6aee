#!/bin/bash
# multi-GPU round trip (run under gpurun --gpus N): tiled parity tests, then bench lines at N = 1..NG
mkdir -p gpurun_out
NG=${NG:-2}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "tiled" 2>&1 | tail -12 | tee gpurun_out/pytest_tiled.log
for N in ${NS:-1 $NG}; do
  if [ "$N" = "1" ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"; fi
  timeout 900 $L bench.py --gpus $N --steps 5 --config ${CFG:-C3} --no-cpu-baseline ${BENCH_ARGS:---no-c5} 2>gpurun_out/multi_${CFG:-C3}_g$N.err | tee gpurun_out/multi_${CFG:-C3}_g$N.json | python tools/jline.py "N=$N"
  tail -3 gpurun_out/multi_${CFG:-C3}_g$N.err
done
