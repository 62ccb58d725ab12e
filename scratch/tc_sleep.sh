#!/bin/bash
for ns in 200 500 1000 2000 4000; do
  GR_TC_DEBUG=$((ns*256)) timeout 600 python bench.py --workload C1 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())['extras']['eval_c4']
print('sleep_ns', $ns, {k:d[k] for k in ('ms','rows_reranked_exactly','lists_identical_to_exact_kernel')})"
done
