#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_submodules.py -x -q 2>&1 | tail -3
out=gpurun_out/r2am_ab.txt
: > $out
run() {
  env "$@" python bench.py --steps 30 --warmup 4 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$*', 'ms_per_step', round(d['ms_per_step'], 3), 'launches', d['gpu_launches'] // d['steps'], 'inter_edges', d['config']['inter_edges_last_iter'])
" >> $out
}
for i in 1 2 3; do
run FB_RA_NQ4=1
run FB_RA_NQ4=0
done
cat $out
