#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2ba_bench_default.json 2>gpurun_out/r2ba_bench_default.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2ba_bench_default.json') if l.startswith('{')][-1]); print('config 2', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('frac_in_step_estimate'), d['roofline_kernels']['gemm_edge']['frac'], d['roofline_kernels']['gemm_edge'].get('frac_in_step_estimate'), d['gpu_launches'], d['cpu_baseline']['value'])
for k,v in d['extras'].items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('error'))"
timeout 600 python bench.py --config 5 --steps 10 --warmup 3 > gpurun_out/r2ba_bench_cfg5.json 2>/dev/null
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2ba_bench_cfg5.json') if l.startswith('{')][-1]); print('config 5', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])"
