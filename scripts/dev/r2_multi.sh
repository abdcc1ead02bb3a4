#!/bin/bash
# N-GPU check of the training step (NCCL all-reduce) and of the default line under torchrun; N = number of visible GPUs
set -x
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config 5 --steps 5 --warmup 3 > gpurun_out/r2u_cfg5_n$N.json 2> gpurun_out/r2u_cfg5_n$N.err; cat gpurun_out/r2u_cfg5_n$N.json; tail -3 gpurun_out/r2u_cfg5_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2u_default_n$N.json 2> gpurun_out/r2u_default_n$N.err; tail -c 2500 gpurun_out/r2u_default_n$N.json; tail -3 gpurun_out/r2u_default_n$N.err
