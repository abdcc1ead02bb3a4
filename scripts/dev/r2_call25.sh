#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r2ac_gpu_tests.txt
timeout 600 python bench.py --config 5 --steps 8 --warmup 3 > gpurun_out/r2ac_bench_cfg5.json 2>/dev/null; python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2ac_bench_cfg5.json') if l.startswith('{')][-1]); print('config 5', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['train'])"
timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('config 2', d['value'], d['ms_per_step'])"
