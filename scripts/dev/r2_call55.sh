#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_model.py -x -q 2>&1 | tail -2
out=gpurun_out/r2ay_ab.txt
: > $out
run() {
  env "$@" python bench.py --steps 30 --warmup 4 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$*', 'ms_per_step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['ms_per_step'],3))
" >> $out
}
for i in 1 2 3; do
run FB_TC5_WHINT=0
run FB_TC5_WHINT=1
done
cat $out
