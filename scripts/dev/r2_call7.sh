#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -5 | tee gpurun_out/r2g_gemm_tests.txt
timeout 300 python scripts/dev/gemm_time.py 2>&1 | tail -14 | tee gpurun_out/r2g_gemm_time.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; tail -c 1200 gpurun_out/r2g_bench.json; tail -3 gpurun_out/r2g_bench.err
