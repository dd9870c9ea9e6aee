#!/bin/bash
# round-end evidence: launch list of one bench step + full ncu capture of the two hot kernels (C3)
mkdir -p gpurun_out
CFG=${CFG:-C3}
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_$CFG.csv python bench.py --config $CFG --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-c5 > gpurun_out/ncu_launch.log 2>&1
timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"sketch_contract|bcd_sweep|knn_kernel|objective_kernel" -c 6 -o gpurun_out/r02_prof_$CFG -f python bench.py --config $CFG --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-c5 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200
