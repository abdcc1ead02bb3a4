#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r2bc_ab.txt
: > $out
for rep in 1 2; do
for v in base rows head both; do
cp variants/gpurun_variants_$v.so fabind_b200/libfabind_b200.so
python bench.py --steps 30 --warmup 4 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', 'ms_per_step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['ms_per_step'],3), 'inter_edges', d['config']['inter_edges_last_iter'])" >> $out
done
done
cat $out
