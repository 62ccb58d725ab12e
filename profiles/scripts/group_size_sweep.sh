#!/bin/bash
for gn in 384 512 768 1024; do
for w in C1 C4; do
  GR_GROUP_NNZ=$gn timeout 600 python bench.py --workload $w --no-cpu --no-e2e --no-extras 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('group_nnz', $gn, '$w', round(d['value']/1e9,2), 'G edges/s', round(d['ms_per_step'],3), 'ms')"
done
done
