#!/bin/bash
# round 2, GPU call 20 (4 GPUs): the fused exchange (k = 2) on a DENSE random state instead of |0...0> (the circuits see
# 450 GB/s, the microbenchmark on |0...0> 660: is the transfer rate data-dependent?), 32 local qubits
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== p2p swap, 4 GPUs, 32 local qubits, dense state"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29670 tools/p2p_swap_probe.py 32 bits dense 2>&1 | grep -v "^\*\|OMP_NUM"
echo "== p2p swap, 4 GPUs, 32 local qubits, |0...0>"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29671 tools/p2p_swap_probe.py 32 bits 2>&1 | grep -v "^\*\|OMP_NUM"
} > gpurun_out/r2_probe20.txt 2>&1
tail -30 gpurun_out/r2_probe20.txt | cut -c1-300
