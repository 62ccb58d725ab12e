#!/bin/bash
mkdir -p gpurun_out
for v in 0 1; do
GR_UO_SERIAL=$v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2951$v bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/r02_uo8_serial$v.json 2> gpurun_out/r02_uo8_serial$v.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_uo8_serial$v.json').read().strip().splitlines()[-1]); print('serial=$v ms/step', d['ms_per_step'], 'per-graph ms', d['roofline']['avg_launch_ms_per_graph'], 'share', d['roofline']['kernel_share_of_step'])"
done
