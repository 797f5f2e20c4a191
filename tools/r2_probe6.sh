#!/bin/bash
# round 2, GPU call 6 (2 GPUs): NVLink peer-memory exchange -- NCCL-worker parity test, sharded bench at 30 and 33 local qubits
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== pytest dist + provider + new kernels"; timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_provider.py -m gpu -x -q 2>&1 | tail -8
echo "== bench --gpus 2 --qubits 30"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 2 --warmup 1 --qubits 30 --no-families > gpurun_out/r2b_bench_2gpu_q31.json 2> gpurun_out/r2b_bench_2gpu_q31.err
echo "rc=$?"; tail -5 gpurun_out/r2b_bench_2gpu_q31.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2b_bench_2gpu_q31.json'))
print({k:d[k] for k in ("value","ms_per_step","schedule")}, d["check"]["parity"], d["e2e"].get("circuit_s"))
PY
echo "== bench --gpus 2 --qubits 33"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2b_bench_2gpu_q34.json 2> gpurun_out/r2b_bench_2gpu_q34.err
echo "rc=$?"; tail -5 gpurun_out/r2b_bench_2gpu_q34.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2b_bench_2gpu_q34.json'))
print({k:d[k] for k in ("value","ms_per_step","schedule","qft","random_circuit")}, d["check"]["parity"], d["e2e"])
PY
} > gpurun_out/r2_probe6.txt 2>&1
tail -40 gpurun_out/r2_probe6.txt | cut -c1-1500
