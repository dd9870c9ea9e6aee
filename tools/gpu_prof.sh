#!/bin/bash
# ncu captures: CFG=C5 KERN=regex
mkdir -p gpurun_out
CFG=${CFG:-C3}
timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"${KERN:-sketch_contract|bcd_sweep}" -c ${COUNT:-2} -o gpurun_out/r02_${TAG:-prof}_$CFG -f python bench.py --config $CFG --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-c5 > gpurun_out/ncu_$CFG.log 2>&1
tail -2 gpurun_out/ncu_$CFG.log | cut -c1-300
