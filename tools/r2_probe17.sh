#!/bin/bash
# round 2, GPU call 17 (1 GPU): partial barriers at stage boundaries (A/B against the full group barrier, flag 2048, debug
# library), gpu suite, QFT probes after the diagonal packing, the bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
DBG=$PWD/qiskit_sdk_py_b200/libqsv_dbg.so
{
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== QV-30 variants (dbg lib): label ms_per_step launch_ms passes frac sm_mhz reasons"
QSV_LIB_PATH=$DBG bash tools/b30_variants.sh "partial A=1" "full_barrier QSV_DEBUG_FLAGS=2048" "partial_again A=1" "full_barrier_again QSV_DEBUG_FLAGS=2048"
echo "== qft probe 33"; timeout 600 python tools/qft_probe.py 33
echo "== qft probe 30"; timeout 600 python tools/qft_probe.py 30
echo "== bench (33 qubits)"; timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r2k_bench_q33.json 2> gpurun_out/r2k_bench_q33.err
echo "rc=$?"; tail -2 gpurun_out/r2k_bench_q33.err; python - <<PY
import json
d=json.load(open('gpurun_out/r2k_bench_q33.json'))
print({k:d[k] for k in ("value","ms_per_step","roofline","e2e")})
print("qft", d["qft"]["pass_section_ms"], d["qft"]["roofline"]["per_pass_ms"])
print("random", d["random_circuit"]["pass_section_ms"], d["random_circuit"]["passes"])
PY
} > gpurun_out/r2_probe17.txt 2>&1
tail -70 gpurun_out/r2_probe17.txt | cut -c1-1500
