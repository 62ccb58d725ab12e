#!/bin/bash
# ncu launch lists (gpu__time_duration.sum) of eager training steps of every model family at its BASELINE shape
mkdir -p gpurun_out
for c in lightgcn_c1 gs_c2 ngcf_c3 gat_c3; do
  python profiles/scripts/r02_model_step.py $c 6 > gpurun_out/r02_step_$c.plain.log 2>&1 || { echo "plain run of $c failed"; tail -5 gpurun_out/r02_step_$c.plain.log; continue; }
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_step_launches_$c.csv \
      python profiles/scripts/r02_model_step.py $c 6 > gpurun_out/r02_step_$c.ncu.log 2>&1
  echo "$c: $(wc -l < gpurun_out/r02_step_launches_$c.csv) csv lines"
done
