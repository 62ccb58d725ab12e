#!/bin/bash
# 2 GPUs: bit-identity of the exact row partition, 1e-5 parity of the user-owner mode, then C5/8 in both modes
mkdir -p gpurun_out
for shape in C1 C4; do
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py $shape 2>&1 | grep -v "^W\|Warning\|\*\*\*\|OMP_NUM" | tail -6
done
for part in user-owner rows; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload C5/8 --steps 5 --warmup 3 --partition $part --no-cpu > gpurun_out/r02_bench_c58_2gpu_$part.json 2> gpurun_out/r02_bench_c58_2gpu_$part.err
grep -v "Warn\|warn\|^W\|OMP\|\*\*\*" gpurun_out/r02_bench_c58_2gpu_$part.err | tail -5
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_c58_2gpu_$part.json").read().strip().splitlines()[-1])
    print("$part", "ms/step", d["ms_per_step"], "value", d["value"], "e2e ms", d["e2e"]["ms_per_step"] if d.get("e2e") else None, "extras", json.dumps(d.get("extras"))[:1500])
except Exception as e:
    print("$part failed", e)
PY
done
