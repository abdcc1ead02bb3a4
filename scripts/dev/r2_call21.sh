#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r2y_gpu_tests.txt
for c in 1 3 4; do timeout 600 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2y_bench_cfg$c.json 2>/dev/null; python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2y_bench_cfg$c.json') if l.startswith('{')][-1]); print('config $c', d['value'], d['unit'], d['ms_per_step'], d['gpu_launches']/5)"; done
timeout 300 python scripts/bench_configs.py 2>&1 | tail -12 | tee gpurun_out/r2y_bench_configs.txt
