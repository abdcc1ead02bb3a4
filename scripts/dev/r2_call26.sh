#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/r2ad_gpu_tests.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2ad_bench_default.json 2>gpurun_out/r2ad_bench_default.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2ad_bench_default.json') if l.startswith('{')][-1]); print('config 2', d['value'], d['ms_per_step'], d['e2e'], d['roofline'], d.get('gpu_launches'))"
