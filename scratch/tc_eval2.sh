#!/bin/bash
for sp in 1 2 3 4; do
  GR_TC_SPLITS=$sp timeout 600 python bench.py --workload C1 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())['extras']['eval_c4']
print('splits', $sp, {k:d[k] for k in ('ms','rows_reranked_exactly','lists_identical_to_exact_kernel')})"
done
