#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_submodules.py tests/test_gpu_plus.py tests/test_gpu_l2.py tests/test_gpu_dataloader_layout.py -x -q 2>&1 | tail -3
for i in 1 2 3; do
python bench.py --steps 30 --warmup 4 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms_per_step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['ms_per_step'],3), 'inter_edges', d['config']['inter_edges_last_iter'])"
done
