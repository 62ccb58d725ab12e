#!/bin/bash
for thr in 256 512 768; do
for w in C4 C1; do
  GR_LONG_ROW_THRESHOLD=$thr timeout 600 python bench.py --workload $w --no-cpu --no-e2e --no-extras 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('long_threshold', $thr, '$w', round(d['value']/1e9,2), 'G edges/s', round(d['ms_per_step'],3), 'ms')"
done
done
