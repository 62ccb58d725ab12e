#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_train_eval.py -m gpu -x -q > gpurun_out/tc_tests.log 2>&1
tail -4 gpurun_out/tc_tests.log
for kp in 0; do
  GR_TC_KPRIME=$kp timeout 600 python bench.py --workload C1 --no-cpu --no-e2e > gpurun_out/tc_bench_$kp.log 2>&1
  tail -1 gpurun_out/tc_bench_$kp.log | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())['extras']['eval_c4']
    print('kprime', $kp, {k:d[k] for k in ('ms','users_per_s','rows_reranked_exactly','exact_only_ms','lists_identical_to_exact_kernel')})
except Exception as e:
    print('kprime', $kp, 'failed', e)"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"topk|max_row_norm" -c 40 --csv --log-file gpurun_out/tc_launches.csv python bench.py --workload C1 --no-cpu --no-e2e > gpurun_out/tc_ncu.log 2>&1
