#!/bin/bash
mkdir -p gpurun_out
for g in 16 8 4 16; do FB_GN_G=$g timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2aa_gn$g.json 2>/dev/null; python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2aa_gn$g.json') if l.startswith('{')][-1]); print('G=$g', round(d['ms_per_step'],3), d['stage_ms_per_step']['edge_elementwise'])"; done
