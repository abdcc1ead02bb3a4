#!/bin/bash
# second GPU call of round 2: full GPU suite, the default bench line (+ extras), config 5 as the main line, training-step phase probe
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 | tee gpurun_out/r2b_gpu_tests.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -c 3000 gpurun_out/r2b_bench.json; tail -5 gpurun_out/r2b_bench.err
timeout 300 python bench.py --config 5 --steps 5 --warmup 3 > gpurun_out/r2b_bench_cfg5.json 2> gpurun_out/r2b_bench_cfg5.err; cat gpurun_out/r2b_bench_cfg5.json; tail -5 gpurun_out/r2b_bench_cfg5.err
timeout 300 python scripts/dev/train_profile.py 2>&1 | tee gpurun_out/r2b_train_phases.txt
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_train_launches.csv python scripts/dev/train_profile.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r2b_train_launches.csv')) if len(r) > 10]
hdr = rows[0]; ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try: v = float(r[vi].replace(',', ''))
    except ValueError: continue
    agg[r[ki][:90]][0] += 1; agg[r[ki][:90]][1] += v
tot = sum(v[1] for v in agg.values())
print('kernels', sum(v[0] for v in agg.values()), 'total_us', tot / 1e3)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f'{v[1]/1e3:10.1f} us {v[0]:6d} x {100*v[1]/tot:5.1f}%  {k}')
PY
