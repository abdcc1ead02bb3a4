#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_misc.py tests/test_gpu_train_forward.py tests/test_gpu_train_reverse.py tests/test_gpu_train_reverse_att.py -x -q 2>&1 | tail -3 | tee gpurun_out/r2aq_train_tests.txt
for i in 1 2; do
timeout 600 python bench.py --config 5 --steps 10 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('config 5', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])" | tee -a gpurun_out/r2aq_cfg5.txt
done
timeout 600 python scripts/dev/train_profile.py 2>/dev/null | tr -d '\n ' | tee gpurun_out/r2aq_train_phases.txt; echo
