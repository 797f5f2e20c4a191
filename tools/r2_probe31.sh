#!/bin/bash
# round 2, GPU call 31 (1 GPU): the full gpu suite and smoke() on the final code of the round
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== pytest gpu"; timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== smoke"; timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
} > gpurun_out/r2_probe31.txt 2>&1
tail -8 gpurun_out/r2_probe31.txt | cut -c1-400
