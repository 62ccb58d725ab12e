#!/bin/bash
# ncu time + DRAM / L2 bytes of the secondary kernels (fused BPR, clip+Adam, rowmap fwd/bwd, GAT aggregate fwd/bwd,
# radix-sort scatter of the graph build) inside eager training steps at the BASELINE shapes
mkdir -p gpurun_out
for c in gat_c3 ngcf_c3; do
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none \
  -k regex:"bpr_fused|opt_adam|opt_sumsq|rowmap|gat_aggregate_kernel|gat_bwd_row_kernel|gat_bwd_col_kernel|rs_scatter|layer_combine" -c 70 \
  --csv --log-file gpurun_out/r02_kernel_bytes_$c.csv python profiles/scripts/r02_model_step.py $c 3 > gpurun_out/r02_kernel_bytes_$c.log 2>&1
echo "$c: $(wc -l < gpurun_out/r02_kernel_bytes_$c.csv) csv lines"
done
