#!/bin/bash
# round 2, GPU call 5 (1 GPU): kernel tests after the CX / run-table / scheduler changes, QFT + random probes at 33
# qubits, full bench line at 33 qubits
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== pytest gpu kernels + golden"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -5
echo "== qft probe 33"; timeout 300 python tools/qft_probe.py 33
echo "== random probe 33"; timeout 300 python tools/random_probe.py 33 20
echo "== bench N=1"; timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_bench_q33.json 2> gpurun_out/r2a_bench_q33.err; echo "rc=$?"; tail -3 gpurun_out/r2a_bench_q33.err; cat gpurun_out/r2a_bench_q33.json
} > gpurun_out/r2_probe5.txt 2>&1
tail -30 gpurun_out/r2_probe5.txt | cut -c1-2500
