#!/bin/bash
# round 2, GPU call 32 (1 GPU): ncu --set full of the committed swap_pairs_kernel (byte lookup tables) at n = 30
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 150 ncu --set full --clock-control none -k regex:swap_pairs_kernel -c 1 -o gpurun_out/r2g_swappairs_n30 python tools/qft_probe.py 30 > gpurun_out/r2g_ncu_swappairs.log 2>&1; echo "rc=$?"
ncu -i gpurun_out/r2g_swappairs_n30.ncu-rep --page raw --csv > gpurun_out/r2g_swappairs_n30_raw.csv 2>/dev/null
rm -f gpurun_out/r2g_swappairs_n30.ncu-rep
} > gpurun_out/r2_probe32.txt 2>&1
tail -3 gpurun_out/r2_probe32.txt
