#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 200 python scripts/dev/tc4_stalls.py 2>&1 | tail -5 | tee gpurun_out/r2q_tc4_stalls.txt
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -8 | tee gpurun_out/r2q_gemm_tests.txt
timeout 300 python scripts/dev/gemm_time.py 2>&1 | tail -14 | tee gpurun_out/r2q_gemm_time.txt
