#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r2aw_sanitizer.txt
echo "== memcheck:" > $out
timeout 1500 compute-sanitizer --tool memcheck python scripts/dev/sanitize_small.py 2>&1 | grep -v "^$" | tail -25 >> $out
echo "== synccheck:" >> $out
timeout 1200 compute-sanitizer --tool synccheck python scripts/dev/sanitize_small.py 2>&1 | grep -v "^$" | tail -8 >> $out
cat $out | cut -c1-300
