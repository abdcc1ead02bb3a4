#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_attention_tc.py -x -q 2>&1 | tail -15 | tee gpurun_out/r2e_attention_tests.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; tail -c 1500 gpurun_out/r2e_bench.json; tail -5 gpurun_out/r2e_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:row_attention_tc -c 2 -o gpurun_out/r2e_xatt python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2>&1
ncu -i gpurun_out/r2e_xatt.ncu-rep --page details 2>/dev/null | grep -E "row_attention_tc|Duration|Registers Per|Executed Ipc|Block Size|Grid Size|Warp Cycles Per Issued" | head -40 | tee gpurun_out/r2e_xatt_summary.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/r2e_gpu_tests.txt
