#!/bin/bash
# First GPU call of round 2 (training path): the gated parity tests, then the config-5 side benchmark in both GEMM modes.
#   gpurun --timeout 600 -- 'bash scripts/dev/round2_first_call.sh'
set -x
FB_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_train_forward.py -q -m gpu 2>&1 | tail -25 | tee gpurun_out/train_forward_tests.txt
timeout 120 python scripts/bench_train.py --steps 2 --warmup 1 2>&1 | tail -2 | tee gpurun_out/bench_train_fp32.json
timeout 120 python scripts/bench_train.py --steps 3 --warmup 1 --gemm bf16 --precision bf16 2>&1 | tail -2 | tee gpurun_out/bench_train_bf16.json
