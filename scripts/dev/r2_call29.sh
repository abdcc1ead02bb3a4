#!/bin/bash
# 2 GPUs: config 5 with the overlapped all-reduce, then with the tail collective; GPU test of the hinted layout on one of them
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dataloader_layout.py -x -q 2>&1 | tail -5 | tee gpurun_out/r2ag_tests.txt
for mode in overlap tail; do
  flag=""; [ $mode = tail ] && flag="--no-overlap"
  NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config 5 --steps 8 --warmup 3 $flag > gpurun_out/r2ag_cfg5_n2_$mode.json 2> gpurun_out/r2ag_cfg5_n2_$mode.err
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2ag_cfg5_n2_$mode.json') if l.startswith('{')][-1]); print('$mode', d['value'], d['ms_per_step'], d['train'])"
done
tail -5 gpurun_out/r2ag_cfg5_n2_overlap.err
