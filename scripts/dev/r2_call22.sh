#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck python scripts/dev/sanitize_small.py 2>&1 | tail -15 | tee gpurun_out/r2z_sanitizer_memcheck.txt
timeout 1200 compute-sanitizer --tool racecheck python scripts/dev/sanitize_small.py 2>&1 | grep -E "Race reported|RACECHECK SUMMARY|hazards\]|done" | sort | uniq -c | head -20 | tee gpurun_out/r2z_sanitizer_racecheck.txt
timeout 900 compute-sanitizer --tool synccheck python scripts/dev/sanitize_small.py 2>&1 | tail -4 | tee gpurun_out/r2z_sanitizer_synccheck.txt
