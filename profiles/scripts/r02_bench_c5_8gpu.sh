#!/bin/bash
# 8 GPUs, C5, the default (user-owner) partition; then the exact row partition without e2e/extras for comparison
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu > gpurun_out/r02_bench_c5_8gpu_user-owner.json 2> gpurun_out/r02_bench_c5_8gpu_user-owner.err
grep -v "Warn\|warn\|^W\|OMP\|\*\*\*" gpurun_out/r02_bench_c5_8gpu_user-owner.err | tail -5
tail -c 4000 gpurun_out/r02_bench_c5_8gpu_user-owner.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu --no-e2e --no-extras --partition rows > gpurun_out/r02_bench_c5_8gpu_rows.json 2> gpurun_out/r02_bench_c5_8gpu_rows.err
grep -v "Warn\|warn\|^W\|OMP\|\*\*\*" gpurun_out/r02_bench_c5_8gpu_rows.err | tail -5
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_c5_8gpu_rows.json').read().strip().splitlines()[-1]); print('rows ms/step', d['ms_per_step'], 'value', d['value'])"
