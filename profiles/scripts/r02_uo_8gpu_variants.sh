#!/bin/bash
mkdir -p gpurun_out
i=0
for v in "GR_UO_COPY_CTAS=16" "GR_UO_COPY_CTAS=24"; do
i=$((i+1))
env $v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2952$i bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/r02_uo8_v$i.json 2> gpurun_out/r02_uo8_v$i.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_uo8_v$i.json').read().strip().splitlines()[-1]); print('$v ms/step', d['ms_per_step'], 'per-graph ms', d['roofline']['avg_launch_ms_per_graph'], 'share', d['roofline']['kernel_share_of_step'])"
done
