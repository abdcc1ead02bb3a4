#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -8 | tee gpurun_out/r2o_gemm_tests.txt
timeout 300 python scripts/dev/gemm_time.py 2>&1 | tail -14 | tee gpurun_out/r2o_gemm_time.txt
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/r2o_gpu_tests.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; head -c 500 gpurun_out/r2o_bench.json; echo; tail -c 500 gpurun_out/r2o_bench.json; tail -3 gpurun_out/r2o_bench.err
