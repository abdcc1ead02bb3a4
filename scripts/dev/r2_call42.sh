#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/dev/train_profile.py 2>/dev/null | tr -d '\n ' | tee gpurun_out/r2ar_train_phases.txt; echo
