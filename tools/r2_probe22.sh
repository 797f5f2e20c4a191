#!/bin/bash
# round 2, GPU call 22 (1 GPU): qsv_swap_qubit_pairs (swaps taken out of the gate stream), full gpu suite, smoke, QFT probes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== qft probe 33"; timeout 600 python tools/qft_probe.py 33
echo "== qft probe 30"; timeout 600 python tools/qft_probe.py 30
echo "== bench (33 qubits)"; timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_bench_q33.json 2> gpurun_out/r2m_bench_q33.err
echo "rc=$?"; tail -2 gpurun_out/r2m_bench_q33.err; python - <<PY
import json
d=json.load(open('gpurun_out/r2m_bench_q33.json'))
print({k:d[k] for k in ("value","ms_per_step","e2e")}, d["roofline"]["frac"], d["roofline"]["launch_ms"])
print("qft", d["qft"]["pass_section_ms"], d["qft"]["roofline"]["per_pass_ms"], d["qft"]["roofline"].get("per_pass_kind"))
print("random", d["random_circuit"]["pass_section_ms"], d["random_circuit"]["passes"])
PY
} > gpurun_out/r2_probe22.txt 2>&1
tail -70 gpurun_out/r2_probe22.txt | cut -c1-1500
