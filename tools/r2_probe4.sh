#!/bin/bash
# round 2, GPU call 4 (2 GPUs): full gpu test suite incl. the 2-rank NCCL test and the plugin tests; sharded bench with
# the golden parity gate at 30 local qubits
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== bench --gpus 2 --qubits 30"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 2 --warmup 1 --qubits 30 > gpurun_out/r2_bench_2gpu_q31.json 2> gpurun_out/r2_bench_2gpu_q31.err
echo "rc=$?"; tail -5 gpurun_out/r2_bench_2gpu_q31.err; cat gpurun_out/r2_bench_2gpu_q31.json
} > gpurun_out/r2_probe4.txt 2>&1
tail -40 gpurun_out/r2_probe4.txt
