#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --config 5 --steps 10 --warmup 3 > gpurun_out/r2bg_bench_cfg5_n2.json 2>/dev/null
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2bg_bench_cfg5_n2.json') if l.startswith('{')][-1]); print('N=2 config 5', d['value'], d['ms_per_step'], d['train']['ms_allreduce_exposed'], d['train']['ms_allreduce_alone'], d['train']['allreduce_alone_busbw_gbs'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 10 --warmup 3 --no-extras > gpurun_out/r2bg_bench_default_n2.json 2>/dev/null
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2bg_bench_default_n2.json') if l.startswith('{')][-1]); print('N=2 config 2', d['value'], d['ms_per_step'], d['e2e']['value'])"
