#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max --clock-control none -k regex:spmm -c 12 --csv --log-file gpurun_out/c4_launches.csv python bench.py --workload ${W:-C4} --no-cpu --no-e2e --no-extras --steps 2 --warmup 1 > gpurun_out/c4_ncu.log 2>&1
echo rc=$?
