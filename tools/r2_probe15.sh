#!/bin/bash
# round 2, GPU call 15 (2 GPUs): peer-memory swap kernel, grid-stride vs one contiguous slice per CTA, at 30 and 33 local
# qubits (the in-circuit exchanges at 33 local qubits run at 410-540 GB/s, the 30-qubit microbenchmark at 688); the 2-GPU
# parity tests and bench after the planner / carry-over changes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for nl in 30 33; do
echo "== p2p swap probe, $nl local qubits"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29660 tools/p2p_swap_probe.py $nl quick 2>&1 | grep -v "^\*\|OMP_NUM"
done
echo "== pytest tests/test_gpu_dist.py"; timeout 900 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -3
echo "== bench --gpus 2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29661 bench.py --gpus 2 --steps 2 --warmup 1 --no-families --no-e2e --no-parity > gpurun_out/r2i_bench_2gpu.json 2> gpurun_out/r2i_bench_2gpu.err
echo "rc=$?"; tail -2 gpurun_out/r2i_bench_2gpu.err | cut -c1-300; python - <<PY
import json
d=json.load(open('gpurun_out/r2i_bench_2gpu.json'))
print({k:d[k] for k in ("value","ms_per_step","schedule")})
PY
} > gpurun_out/r2_probe15.txt 2>&1
tail -60 gpurun_out/r2_probe15.txt | cut -c1-1200
