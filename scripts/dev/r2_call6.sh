#!/bin/bash
# ncu full captures (with source counters) of the edge-level tc4 GEMM and the node-level tc3 GEMM inside the bench step
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc4 -s 40 -c 2 -o gpurun_out/r2f_tc4 python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc3 -s 100 -c 3 -o gpurun_out/r2f_tc3 python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 300 python scripts/dev/gemm_probe2.py 2>&1 | tail -8 | tee gpurun_out/r2f_gemm_probe2.txt
