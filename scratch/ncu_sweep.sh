#!/bin/bash
# full ncu capture of the sweep kernel(s) on one config: scratch/ncu_sweep.sh C3 [variant]
CFG=${1:-C3}; V=${2:-0}
mkdir -p gpurun_out
FDB_SWEEP_VARIANT=$V timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"bcd_sweep" -s 20 -c 2 -f -o gpurun_out/sweep_${CFG}_v$V python bench.py --config $CFG --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_sweep.log 2>&1
tail -3 gpurun_out/ncu_sweep.log
ls -la gpurun_out/*.ncu-rep
