#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc4 -s 41 -c 2 -o gpurun_out/r2r_tc4 python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc5 -s 60 -c 4 -o gpurun_out/r2r_tc5 python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2r_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2r_bench_under_ncu.json 2>/dev/null
python scripts/summarize_launches.py gpurun_out/r2r_launches.csv > gpurun_out/r2r_launches.txt 2>&1; head -30 gpurun_out/r2r_launches.txt
ls -la gpurun_out/r2r*
