#!/bin/bash
# round 2, one B200: GPU test suite, the default bench line, its ncu launch list, ncu full capture of the nomination
# kernel (final version) and the launch list of the evaluation kernels
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r02_final_tests.log 2>&1; tail -2 gpurun_out/r02_final_tests.log
timeout 1200 python bench.py > gpurun_out/r02_bench_c5_1gpu.json 2> gpurun_out/r02_bench_c5_1gpu.err || { echo "bench failed"; tail -5 gpurun_out/r02_bench_c5_1gpu.err; exit 1; }
tail -c 300 gpurun_out/r02_bench_c5_1gpu.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_c5.csv python bench.py --no-e2e --no-cpu --no-extras > gpurun_out/r02_ncu_launches.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r02_launches_c5.csv
timeout 250 ncu --set full --clock-control none --import-source on -k regex:topk_tc_candidates -s 2 -c 1 -o gpurun_out/r02_topk_tc_full -f python profiles/scripts/r02_eval_c4.py 1 > gpurun_out/r02_topk_tc_full.log 2>&1
echo "topk full rc=$?"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"topk|rescore|max_row_norm" --csv --log-file gpurun_out/r02_launches_eval_c4.csv python profiles/scripts/r02_eval_c4.py 1 > gpurun_out/r02_eval_c4_ncu.log 2>&1
echo "eval launch list rc=$?"
