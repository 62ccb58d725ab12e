#!/bin/bash
mkdir -p gpurun_out
timeout 800 ncu --set full --clock-control none --import-source on -k regex:topk_tc_candidates -s 1 -c 1 -f -o gpurun_out/r01_topk_tc_full python bench.py --workload C1 --no-cpu --no-e2e > gpurun_out/tc_ncu_full.log 2>&1
echo rc=$?; ls -la gpurun_out/r01_topk_tc_full.ncu-rep
