#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2bb_train_launches.csv python scripts/dev/train_profile.py > gpurun_out/r2bb_train_phases.txt 2> gpurun_out/r2bb_train_profile.err
python scripts/summarize_launches.py gpurun_out/r2bb_train_launches.csv > gpurun_out/r2bb_train_launches.txt
head -50 gpurun_out/r2bb_train_launches.txt
cat gpurun_out/r2bb_train_phases.txt | tr -d '\n ' | head -c 1500
