#!/bin/bash
# In-situ cost of each launch category: step time with the category's launches dropped (results are garbage;
# timing only).  Categories: 0 gemm_edge 1 gemm_node 2 gemm_pair 3 gemm_pair0 4 edge_elementwise 5 attention 6 graph_misc
out=gpurun_out/skip_probe.txt
: > $out
for m in 0 1 2 4 16 32 3 63; do
  FB_SKIP_CATS=$m python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('mask', $m, 'ms_per_step', round(d['ms_per_step'], 3), 'launches', d['gpu_launches'], 'inter_edges', d['config']['inter_edges_last_iter'])
" >> $out
done
cat $out
