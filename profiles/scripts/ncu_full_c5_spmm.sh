#!/bin/bash
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:spmm_ -c 3 -f -o gpurun_out/r01_c5_spmm_full \
  python bench.py --workload C5 --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras > gpurun_out/ncu_c5.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/r01_c5_spmm_full.ncu-rep
