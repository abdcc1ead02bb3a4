#!/bin/bash
# ncu --set full of the HBM / L2-side kernels of the final build (segment reduce, gather-add, attention, geometry), mid-forward
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:'gcl_edge_pre|gcl_node|inter_logit|inter_aggregate|row_attention|pair_gather|radial_kernel|las_step' -s 160 -c 20 -o gpurun_out/r2ah_hbm python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2> gpurun_out/r2ah_ncu.err
ls -la gpurun_out/r2ah_hbm.ncu-rep
ncu -i gpurun_out/r2ah_hbm.ncu-rep --page raw --csv > gpurun_out/r2ah_hbm_raw.csv 2>/dev/null
wc -c gpurun_out/r2ah_hbm_raw.csv
