#!/bin/bash
# one GPU round: tests, C3 bench, launch list, full ncu capture of the two main kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 120 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --config C3 --steps 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 2500 gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c3.csv python bench.py --config C3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bcd_sweep|sketch_contract|knn_kernel" -s 14 -c 6 -f -o gpurun_out/prof_c3 python bench.py --config C3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/
