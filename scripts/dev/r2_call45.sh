#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r2as_gpu_tests.txt
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2as_bench_default.json 2>gpurun_out/r2as_bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2as_bench_reference.json 2>/dev/null
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/r2as_smoke.txt
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2as_bench_default.json') if l.startswith('{')][-1]); print('config 2', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline_kernels']['gemm_edge']['frac'], d['gpu_launches'])
for k,v in d['extras'].items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('error'))
r=json.loads([l for l in open('gpurun_out/r2as_bench_reference.json') if l.startswith('{')][-1]); print('reference', r['value'], r.get('cpu_baseline'))"
