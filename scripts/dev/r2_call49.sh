#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc5' -s 80 -c 6 -o gpurun_out/r2av_tc5 python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2> gpurun_out/r2av_ncu.err
ls -la gpurun_out/r2av_tc5.ncu-rep
