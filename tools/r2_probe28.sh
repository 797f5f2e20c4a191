#!/bin/bash
# round 2, GPU call 28 (4 GPUs): the driver N = 4 command on the final code (parity gate, families, e2e)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== bench --gpus 4"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29661 bench.py --gpus 4 --steps 2 --warmup 1 > gpurun_out/r2s_bench_4gpu.json 2> gpurun_out/r2s_bench_4gpu.err
echo "rc=$?"; tail -2 gpurun_out/r2s_bench_4gpu.err | cut -c1-300; python - <<PY
import json
d=json.load(open('gpurun_out/r2s_bench_4gpu.json'))
print({k:d[k] for k in ("value","ms_per_step","schedule","e2e","qft","random_circuit")}, d["check"]["parity"])
PY
} > gpurun_out/r2_probe28.txt 2>&1
tail -40 gpurun_out/r2_probe28.txt | cut -c1-2500
