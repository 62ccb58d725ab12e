#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"topk|max_row_norm" -c 40 --csv --log-file gpurun_out/tc_launches.csv python bench.py --workload C1 --no-cpu --no-e2e > gpurun_out/tc_ncu.log 2>&1
echo rc=$?
GR_TC_DEBUG=1 timeout 600 python bench.py --workload C1 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())['extras']['eval_c4']; print('debug1 (no selection) ms', d['ms'], d['rows_reranked_exactly'])"
