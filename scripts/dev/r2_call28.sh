#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dataloader_layout.py tests/test_gpu_model.py tests/test_gpu_edges.py -x -q 2>&1 | tail -25 | tee gpurun_out/r2af_tests.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2af_bench.json 2>gpurun_out/r2af_bench.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2af_bench.json') if l.startswith('{')][-1]); print('config 2', d['value'], d['ms_per_step'], d['e2e'], d['extras'].get('e2e_prepared'))"
