#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2l_launches.csv python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2>&1
python scripts/summarize_launches.py gpurun_out/r2l_launches.csv > gpurun_out/r2l_launches.txt 2>&1; head -50 gpurun_out/r2l_launches.txt
