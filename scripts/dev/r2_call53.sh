#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('default: config 2', round(d['ms_per_step'],3), 'train_step extra', d['extras']['train_step'].get('ms_per_step'), d['extras']['train_step'].get('ms_forward_backward'))"
done
timeout 600 python bench.py --config 5 --steps 10 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('config 5', d['ms_per_step'], d['e2e']['ms_per_step'])"
