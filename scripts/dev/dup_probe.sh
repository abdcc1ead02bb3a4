#!/bin/bash
# In-situ cost of single kernels (diagnostic build: python -m fabind_b200.build --force --diag): the masked kernels are launched twice
# (FB_KDUP, forward.cu::stage), results intact; FB_GN_MODE bit0 = gcl_node without its cooperative pass for the high-degree rows.
out=gpurun_out/dup_probe.txt
: > $out
run() {
  env "$@" python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$*', 'ms_per_step', round(d['ms_per_step'], 3), 'launches', d['gpu_launches'] // d['steps'], 'inter_edges', d['config']['inter_edges_last_iter'])
" >> $out
}
run FB_KDUP=0
run FB_KDUP=1
run FB_KDUP=2
run FB_KDUP=4
run FB_KDUP=8
run FB_KDUP=16
run FB_KDUP=32
run FB_KDUP=64
run FB_KDUP=0
run FB_GN_MODE=1
cat $out
