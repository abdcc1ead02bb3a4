#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r2be_cfg4_launches.csv python bench.py --config 4 --steps 1 --warmup 3 > /dev/null 2>&1
python scripts/summarize_launches.py gpurun_out/r2be_cfg4_launches.csv > gpurun_out/r2be_cfg4_launches.txt; head -40 gpurun_out/r2be_cfg4_launches.txt
