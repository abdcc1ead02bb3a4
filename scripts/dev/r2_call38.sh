#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/dev/train_profile.py > gpurun_out/r2an_train_phases.txt 2>/dev/null
cat gpurun_out/r2an_train_phases.txt | tr -d '\n ' ; echo
