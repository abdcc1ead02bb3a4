#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/dev/tc5_stalls.py 2>&1 | tee gpurun_out/r2az_tc5_role_stalls.txt | cut -c1-700
