#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r2ak_ab.txt
: > $out
run() {
  env "$@" python bench.py --steps 30 --warmup 4 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$*', 'ms_per_step', round(d['ms_per_step'], 3), 'launches', d['gpu_launches'] // d['steps'], 'inter_edges', d['config']['inter_edges_last_iter'])
" >> $out
}
for i in 1 2 3; do
run FB_IL_PAIRU=1
run FB_IL_PAIRU=0
done
run FB_IL_PAIRU=1 FB_KDUP=16
run FB_IL_PAIRU=0 FB_KDUP=16
run FB_IL_PAIRU=1 FB_GN_HEADS=0
cat $out
