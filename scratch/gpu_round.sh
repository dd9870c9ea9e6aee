#!/bin/bash
# one GPU round: tests, smoke, C3 bench, launch list, full ncu capture of the main kernels
CFG=${1:-C3}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 200 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --config $CFG --steps 5 > gpurun_out/bench_$CFG.json 2> gpurun_out/bench_$CFG.err; tail -c 3000 gpurun_out/bench_$CFG.json; tail -3 gpurun_out/bench_$CFG.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$CFG.csv python bench.py --config $CFG --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"bcd_sweep|sketch_contract|knn_kernel" -c 4 -f -o gpurun_out/prof_$CFG python bench.py --config $CFG --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/ | head -20
