#!/bin/bash
mkdir -p gpurun_out
timeout 800 ncu --set full --clock-control none --import-source on -k regex:spmm_long -s 3 -c 1 -f -o gpurun_out/r01_c4_long_full python bench.py --workload C4 --no-cpu --no-e2e --no-extras --steps 2 --warmup 1 > gpurun_out/c4_ncu_full.log 2>&1
echo rc=$?
