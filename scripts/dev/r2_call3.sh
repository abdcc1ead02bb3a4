#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 | tee gpurun_out/r2c_gpu_tests.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -c 2500 gpurun_out/r2c_bench.json; tail -5 gpurun_out/r2c_bench.err
timeout 300 python scripts/dev/train_profile.py 2>&1 | tee gpurun_out/r2c_train_phases.txt
