#!/bin/bash
# round 2, GPU call 8 (2 GPUs): P2P swap sweep, overlapped exchange: parity tests + bench A/B at 33 local qubits
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== pytest dist"; timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -5
echo "== p2p swap sweep (30 local qubits)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/p2p_swap_probe.py 30 2>&1 | grep -v "^W\|Warning\|warn"
for ovl in 1 0; do
echo "== bench --gpus 2 --qubits 33, overlap=$ovl"
QSV_DIST_OVERLAP=$ovl timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2956$ovl bench.py --gpus 2 --steps 2 --warmup 1 --no-families --no-e2e > gpurun_out/r2c_bench_2gpu_q34_ovl$ovl.json 2> gpurun_out/r2c_bench_2gpu_q34_ovl$ovl.err
echo "rc=$?"; tail -3 gpurun_out/r2c_bench_2gpu_q34_ovl$ovl.err; python - <<PY
import json
d=json.load(open('gpurun_out/r2c_bench_2gpu_q34_ovl$ovl.json'))
print({k:d[k] for k in ("value","ms_per_step","schedule")}, d["check"]["parity"])
PY
done
} > gpurun_out/r2_probe8.txt 2>&1
tail -60 gpurun_out/r2_probe8.txt | cut -c1-700
