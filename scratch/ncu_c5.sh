#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r01_bench_c5_default.log 2>&1 || { echo "bench failed"; tail -5 gpurun_out/r01_bench_c5_default.log; exit 1; }
tail -1 gpurun_out/r01_bench_c5_default.log | cut -c1-600
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:spmm_ -c 4 -f -o gpurun_out/r01_c5_spmm_full \
  python bench.py --workload C5 --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras > gpurun_out/ncu_c5.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_c5.log | cut -c1-300; ls -la gpurun_out/*.ncu-rep
