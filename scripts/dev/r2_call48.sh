#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/dev/config4_phases.py 2>/dev/null | tail -2 | tee gpurun_out/r2au_config4_phases.txt
timeout 600 python scripts/dev/config3_phases.py 2>/dev/null | tail -2 | tee gpurun_out/r2au_config3_phases.txt
