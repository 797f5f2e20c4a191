#!/bin/bash
# round 2, GPU call 7 (1 GPU): producer-wait A/B, per-warp profile, dense-k probe, same-config timing, ncu captures
cd "$(dirname "$0")/.."
DBG=$PWD/qiskit_sdk_py_b200/libqsv_dbg.so
mkdir -p gpurun_out
{
echo "== pytest gpu kernels"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -4
echo "== QV-30 variants (dbg lib): parked producer wait (default) vs spinning (1024)"
QSV_LIB_PATH=$DBG bash tools/b30_variants.sh "parked A=1" "spin QSV_DEBUG_FLAGS=1024" "parked_again A=1"
echo "== phase profile, all warps"; QSV_LIB_PATH=$DBG timeout 300 python tools/prof_dump.py 30 10
echo "== dense-k probe"; timeout 300 python tools/dense_k_probe.py 30
echo "== same config"; timeout 900 python tools/same_config.py 2>&1 | grep -v Warning
echo "== ncu launch list (bench --qubits 28 --depth 8)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_q28.csv python bench.py --qubits 28 --depth 8 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-families > gpurun_out/r2_ncu_list.log 2>&1; echo "rc=$?"
echo "== ncu full, pass kernel at n=30"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_pass_kernel -s 40 -c 2 -o gpurun_out/r2_tilepass_n30 python bench.py --qubits 30 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-families > gpurun_out/r2_ncu_full.log 2>&1; echo "rc=$?"
ncu -i gpurun_out/r2_tilepass_n30.ncu-rep --page raw --csv > gpurun_out/r2_tilepass_n30_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_tilepass_n30.ncu-rep --page source --csv > gpurun_out/r2_tilepass_n30_source.csv 2>/dev/null
ls -la gpurun_out/r2_tilepass_n30*
echo "== ncu full, dense-k kernel k=5,6 at n=28"
timeout 600 ncu --set full --clock-control none -k regex:dense_k_dmma -c 12 -o gpurun_out/r2_densek_n28 python tools/dense_k_probe.py 28 > gpurun_out/r2_ncu_densek.log 2>&1; echo "rc=$?"
ncu -i gpurun_out/r2_densek_n28.ncu-rep --page raw --csv > gpurun_out/r2_densek_n28_raw.csv 2>/dev/null
rm -f gpurun_out/r2_densek_n28.ncu-rep
} > gpurun_out/r2_probe7.txt 2>&1
tail -60 gpurun_out/r2_probe7.txt | cut -c1-400
