#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc4 -s 6 -c 1 -o gpurun_out/r2h_tc4_e2 python scripts/dev/gemm_time.py > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc4 -s 50 -c 1 -o gpurun_out/r2h_tc4_c1 python scripts/dev/gemm_time.py > /dev/null 2>&1
ls -la gpurun_out/r2h*
