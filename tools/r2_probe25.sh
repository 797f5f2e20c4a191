#!/bin/bash
# round 2, GPU call 25 (8 GPUs): the driver N=8 command on the final code of the round (swap relabelling, planner changes)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== bench --gpus 8 (33 local qubits)"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 8 --steps 2 --warmup 1 > gpurun_out/r2p_bench_8gpu_q36.json 2> gpurun_out/r2p_bench_8gpu_q36.err
echo "rc=$?"; tail -4 gpurun_out/r2p_bench_8gpu_q36.err; cat gpurun_out/r2p_bench_8gpu_q36.json
} > gpurun_out/r2_probe25.txt 2>&1
tail -12 gpurun_out/r2_probe25.txt | cut -c1-3000
