#!/bin/bash
# round 2, GPU call 13 (1 GPU): gpu suite after the planner changes (in-tile diagonals behind a tail join the pass; partial
# plans), QFT-33 / random-33 per-pass probes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== qft probe 33"; timeout 600 python tools/qft_probe.py 33
echo "== random probe 33"; timeout 600 python tools/random_probe.py 33
echo "== qft probe 30"; timeout 600 python tools/qft_probe.py 30
} > gpurun_out/r2_probe13.txt 2>&1
tail -60 gpurun_out/r2_probe13.txt | cut -c1-400
