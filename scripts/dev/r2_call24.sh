#!/bin/bash
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/r2ab_smoke.txt
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2ab_bench_default.json 2> gpurun_out/r2ab_bench_default.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2ab_bench_default.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
print('roofline', {k:d['roofline'][k] for k in ('achieved','frac','share_of_step','launches_per_step')})
print('edge', {k:d['roofline_kernels']['gemm_edge'][k] for k in ('achieved','frac','avg_launch_ms','traffic','launches_per_step')})
print('step', d['roofline_step'])
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['parity'])
print({k:(v.get('value'),v.get('ms_per_step')) for k,v in d['extras'].items()})
PY
tail -2 gpurun_out/r2ab_bench_default.err
