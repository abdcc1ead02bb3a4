#!/bin/bash
set -x
mkdir -p gpurun_out
( time timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2t_bench_default.json 2> gpurun_out/r2t_bench_default.err ) 2>&1 | tail -4
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2t_bench_default.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks')})
print('roofline', {k:d['roofline'][k] for k in ('kernel','achieved','frac','share_of_step')} if d.get('roofline') else None)
print('edge', {k:d['roofline_kernels']['gemm_edge'][k] for k in ('achieved','frac','avg_launch_ms','traffic')})
print('cpu', d.get('cpu_baseline'))
print('extras', json.dumps(d.get('extras'))[:1500])
PY
tail -3 gpurun_out/r2t_bench_default.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2t_bench_reference.json 2> gpurun_out/r2t_bench_reference.err ) 2>&1 | tail -4; cat gpurun_out/r2t_bench_reference.json | cut -c1-700
