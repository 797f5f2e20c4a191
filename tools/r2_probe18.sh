#!/bin/bash
# round 2, GPU call 18 (4 GPUs): what limits the fused peer-memory exchange at 4 GPUs (450 GB/s per GPU and direction
# against 688 with one pair active)?  swap (remote load + remote store) vs push only vs pull only, all ranks active, k = 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
DBG=$PWD/qiskit_sdk_py_b200/libqsv_dbg.so
{
echo "== p2p modes, 4 GPUs, 32 local qubits"
QSV_LIB_PATH=$DBG timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29670 tools/p2p_swap_probe.py 32 modes 2>&1 | grep -v "^\*\|OMP_NUM"
echo "== nvidia-smi topo"; nvidia-smi topo -m 2>&1 | head -20
echo "== nvlink status (GPU 0)"; nvidia-smi nvlink -s -i 0 2>&1 | head -24
} > gpurun_out/r2_probe18.txt 2>&1
tail -60 gpurun_out/r2_probe18.txt | cut -c1-300
