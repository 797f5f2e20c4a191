#!/bin/bash
# round 2, GPU call 10 (4 GPUs): full gpu suite (shot batching, dist), XOR-ordered P2P exchange A/B at 33 local qubits
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for ovl in 1 0; do
echo "== bench --gpus 4 --qubits 33, overlap=$ovl"
QSV_DIST_OVERLAP=$ovl timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2958$ovl bench.py --gpus 4 --steps 2 --warmup 1 --no-families --no-e2e --no-parity > gpurun_out/r2e_bench_4gpu_ovl${ovl}.json 2> gpurun_out/r2e_bench_4gpu_ovl${ovl}.err
echo "rc=$?"; tail -2 gpurun_out/r2e_bench_4gpu_ovl${ovl}.err; python - <<PY
import json
d=json.load(open('gpurun_out/r2e_bench_4gpu_ovl${ovl}.json'))
print({k:d[k] for k in ("value","ms_per_step","schedule")})
PY
done
} > gpurun_out/r2_probe10.txt 2>&1
tail -30 gpurun_out/r2_probe10.txt | cut -c1-900
