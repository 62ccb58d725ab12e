#!/bin/bash
# round 2, one B200: the default bench line, its ncu launch list, and ncu dram/lts bytes of one SpMM layer at C1 / C4
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/r02_bench_c5_1gpu.json 2> gpurun_out/r02_bench_c5_1gpu.err || { echo "bench failed"; tail -5 gpurun_out/r02_bench_c5_1gpu.err; exit 1; }
tail -c 600 gpurun_out/r02_bench_c5_1gpu.json; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_c5.csv python bench.py --no-e2e --no-cpu --no-extras > gpurun_out/r02_ncu_launches.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r02_launches_c5.csv
for W in C1 C4; do
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:spmm_ -s 18 -c 6 --csv --log-file gpurun_out/r02_spmm_bytes_$W.csv python bench.py --workload $W --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/r02_ncu_bytes_$W.log 2>&1
echo "$W bytes rc=$?"; wc -l gpurun_out/r02_spmm_bytes_$W.csv
done
