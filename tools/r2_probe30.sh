#!/bin/bash
# round 2, GPU call 30 (1 GPU): swap_pairs_kernel with byte lookup tables instead of per-pair 64-bit shift arithmetic (the first
# version was ALU-bound): its tests, the planner / golden tests that go through it, QFT probes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== pytest (swap pairs, qft, golden)"; timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py -m gpu -x -q -k "swap or qft or golden or stream" 2>&1 | tail -3
echo "== qft probe 33"; timeout 200 python tools/qft_probe.py 33 | tail -3
echo "== qft probe 30"; timeout 100 python tools/qft_probe.py 30 | tail -3
} > gpurun_out/r2_probe30.txt 2>&1
tail -20 gpurun_out/r2_probe30.txt | cut -c1-400
