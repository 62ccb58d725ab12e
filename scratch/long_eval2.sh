#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_models.py -m gpu -x -q 2>&1 | tail -2
for w in C4 C1 C5/8 C5; do
  timeout 900 python bench.py --workload $w --no-cpu --no-e2e --no-extras 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$w', round(d['value']/1e9,2), 'G edges/s', round(d['ms_per_step'],3), 'ms', 'frac', round(d['roofline']['frac'],3), d['clocks'])"
done
