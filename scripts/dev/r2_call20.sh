#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gcl_edge_pre|gcl_node|inter_logit|inter_aggregate|row_attention|pair_gather' -s 90 -c 7 -o gpurun_out/r2x_mid python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/r2x_mid.ncu-rep
