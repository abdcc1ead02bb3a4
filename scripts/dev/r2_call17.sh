#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -4
timeout 300 python scripts/dev/gemm_time.py 2>&1 | head -5 | tee gpurun_out/r2s_gemm_time.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:gemm_tc4 -s 6 -c 1 python scripts/dev/gemm_time.py 2>&1 | grep -E "dram__|gpu__time" | tee gpurun_out/r2s_tc4_dram.txt
