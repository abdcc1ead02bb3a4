#!/bin/bash
mkdir -p gpurun_out
# launch list of the final build (same command as the bench line, fewer steps): one GPU
CUDA_VISIBLE_DEVICES=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r2at_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2>&1
python scripts/summarize_launches.py gpurun_out/r2at_launches.csv > gpurun_out/r2at_launches.txt; head -16 gpurun_out/r2at_launches.txt
# the default line on 2 GPUs (what the driver's scaling run launches)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2at_bench_default_n2.json 2> gpurun_out/r2at_bench_default_n2.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2at_bench_default_n2.json') if l.startswith('{')][-1]); print('N=2 config 2', d['value'], d['ms_per_step'], d['e2e']['value'])
for k,v in d['extras'].items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('allreduce_alone_ms'), v.get('allreduce_alone_busbw_gbs'), v.get('error'))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --config 5 --steps 10 --warmup 3 > gpurun_out/r2at_bench_cfg5_n2.json 2>/dev/null
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2at_bench_cfg5_n2.json') if l.startswith('{')][-1]); print('N=2 config 5', d['value'], d['ms_per_step'], d['train'])"
