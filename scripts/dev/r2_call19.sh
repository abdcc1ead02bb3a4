#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r2w_gpu_tests.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; head -c 420 gpurun_out/r2w_bench.json; echo; tail -2 gpurun_out/r2w_bench.err
