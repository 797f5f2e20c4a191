#!/bin/bash
# round 2, GPU call 26 (2 GPUs): N = 2 after the layout restore keeps the lowest local bits out of the exchange
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== pytest tests/test_gpu_dist.py"; timeout 900 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -3
echo "== bench --gpus 2"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29661 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2q_bench_2gpu.json 2> gpurun_out/r2q_bench_2gpu.err
echo "rc=$?"; tail -2 gpurun_out/r2q_bench_2gpu.err | cut -c1-300; python - <<PY
import json
d=json.load(open('gpurun_out/r2q_bench_2gpu.json'))
print({k:d[k] for k in ("value","ms_per_step","schedule","e2e","qft","random_circuit")}, d["check"]["parity"])
PY
} > gpurun_out/r2_probe26.txt 2>&1
tail -40 gpurun_out/r2_probe26.txt | cut -c1-2500
