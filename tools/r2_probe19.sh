#!/bin/bash
# round 2, GPU call 19 (4 GPUs): the fused exchange (k = 2) at 33 local qubits for several choices of the swapped local
# bits (the microbenchmark reaches 670 GB/s with bits [31, 30] at 32 local qubits, the circuits see 450)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for nl in 33 32; do
echo "== p2p swap, 4 GPUs, $nl local qubits"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29670 tools/p2p_swap_probe.py $nl bits 2>&1 | grep -v "^\*\|OMP_NUM"
done
} > gpurun_out/r2_probe19.txt 2>&1
tail -30 gpurun_out/r2_probe19.txt | cut -c1-300
