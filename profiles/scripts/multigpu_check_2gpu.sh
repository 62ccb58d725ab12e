#!/bin/bash
for shape in C1 C4; do
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py $shape 2>&1 | grep -v "^W\|Warning\|\*\*\*\|OMP_NUM" | tail -4
done
