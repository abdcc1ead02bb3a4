#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/r2k_gpu_tests.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; head -c 600 gpurun_out/r2k_bench.json; echo; tail -c 600 gpurun_out/r2k_bench.json; tail -3 gpurun_out/r2k_bench.err
