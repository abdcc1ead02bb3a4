#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_submodules.py tests/test_gpu_plus.py -x -q 2>&1 | tail -4 | tee gpurun_out/r2aj_tests.txt
out=gpurun_out/r2aj_ab.txt
: > $out
run() {
  env "$@" python bench.py --steps 24 --warmup 4 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$*', 'ms_per_step', round(d['ms_per_step'], 3), 'launches', d['gpu_launches'] // d['steps'], 'inter_edges', d['config']['inter_edges_last_iter'])
" >> $out
}
run FB_IL_MINB=4
run FB_IL_MINB=2
run FB_IL_MINB=3
run FB_IL_MINB=6
run FB_IL_MINB=4
run FB_IL_MINB=4 FB_KDUP=16
run FB_IL_MINB=6 FB_KDUP=16
run FB_IL_MINB=3 FB_KDUP=16
cat $out
