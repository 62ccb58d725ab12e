#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r01_bench_c5_default.log 2>&1 || { echo "bench failed"; tail -5 gpurun_out/r01_bench_c5_default.log; exit 1; }
tail -1 gpurun_out/r01_bench_c5_default.log | cut -c1-300
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_c5.csv python bench.py --no-e2e --no-cpu --no-extras > gpurun_out/ncu_launches.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/r01_launches_c5.csv
