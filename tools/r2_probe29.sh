#!/bin/bash
# round 2, GPU call 29 (1 GPU): ncu captures of the FINAL kernels of the round (each command exited 0 without ncu in tools/r2_probe22.sh /
# r2_probe27.sh): launch list of a short bench, --set full of the pass kernel at n = 30 and of swap_pairs_kernel at n = 30
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== ncu launch list (bench --qubits 28 --depth 8)"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches_bench_q28.csv python bench.py --qubits 28 --depth 8 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2f_ncu_list.log 2>&1; echo "rc=$?"
echo "== ncu full, pass kernel at n=30"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tile_pass_kernel -s 40 -c 2 -o gpurun_out/r2f_tilepass_n30 python bench.py --qubits 30 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-families > gpurun_out/r2f_ncu_full.log 2>&1; echo "rc=$?"
ncu -i gpurun_out/r2f_tilepass_n30.ncu-rep --page raw --csv > gpurun_out/r2f_tilepass_n30_raw.csv 2>/dev/null
ncu -i gpurun_out/r2f_tilepass_n30.ncu-rep --page source --csv > gpurun_out/r2f_tilepass_n30_source.csv 2>/dev/null
rm -f gpurun_out/r2f_tilepass_n30.ncu-rep
echo "== ncu full, swap_pairs_kernel at n=30 (qft probe)"
timeout 300 ncu --set full --clock-control none -k regex:swap_pairs_kernel -c 1 -o gpurun_out/r2f_swappairs_n30 python tools/qft_probe.py 30 > gpurun_out/r2f_ncu_swappairs.log 2>&1; echo "rc=$?"
ncu -i gpurun_out/r2f_swappairs_n30.ncu-rep --page raw --csv > gpurun_out/r2f_swappairs_n30_raw.csv 2>/dev/null
rm -f gpurun_out/r2f_swappairs_n30.ncu-rep
ls -la gpurun_out/r2f_*
} > gpurun_out/r2_probe29.txt 2>&1
tail -30 gpurun_out/r2_probe29.txt | cut -c1-300
