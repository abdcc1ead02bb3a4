#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r2bf_gpu_tests.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/r2bf_smoke.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2bf_bench.json 2>/dev/null
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2bf_bench.json') if l.startswith('{')][-1]); print('config 2', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])
for k,v in d['extras'].items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('error'))"
