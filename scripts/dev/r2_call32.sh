#!/bin/bash
# diag build: correctness of the gcl_node head CTAs / one-wave inter_logit grid, then A/B step times
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_submodules.py tests/test_gpu_plus.py tests/test_gpu_l2.py -x -q 2>&1 | tail -4 | tee gpurun_out/r2ai_tests.txt
out=gpurun_out/r2ai_ab.txt
: > $out
run() {
  env "$@" python bench.py --steps 24 --warmup 4 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$*', 'ms_per_step', round(d['ms_per_step'], 3), 'launches', d['gpu_launches'] // d['steps'], 'inter_edges', d['config']['inter_edges_last_iter'])
" >> $out
}
run FB_GN_HEADS=1 FB_IL_WAVE=1
run FB_GN_HEADS=0 FB_IL_WAVE=0
run FB_GN_HEADS=1 FB_IL_WAVE=0
run FB_GN_HEADS=0 FB_IL_WAVE=1
run FB_GN_HEADS=1 FB_IL_WAVE=1
run FB_GN_HEADS=0 FB_IL_WAVE=0
cat $out
