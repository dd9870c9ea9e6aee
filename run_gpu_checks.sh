#!/bin/bash
# convenience wrapper used with gpurun: runs the GPU test-suite and leaves the log under gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout ${1:-900} python -m pytest tests -m gpu -x -q 2>&1 | tail -60 | tee gpurun_out/pytest_gpu.log
