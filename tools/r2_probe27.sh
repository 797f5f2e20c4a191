#!/bin/bash
# round 2, GPU call 27 (1 GPU): the driver's N = 1 commands on the final code: bench.py (defaults) and the reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== bench.py (defaults)"; timeout 1200 python bench.py > gpurun_out/r2r_bench_q33.json 2> gpurun_out/r2r_bench_q33.err
echo "rc=$?"; tail -2 gpurun_out/r2r_bench_q33.err | cut -c1-300; python - <<PY
import json
d=json.load(open('gpurun_out/r2r_bench_q33.json'))
print({k:d[k] for k in ("value","ms_per_step","roofline","cpu_baseline","e2e","gpu_launches","clocks")})
print("qft", d["qft"]["pass_section_ms"], d["qft"]["roofline"]["per_pass_ms"])
print("random", d["random_circuit"]["pass_section_ms"], d["random_circuit"]["passes"])
PY
echo "== bench.py --impl reference --steps 1 --warmup 0"; timeout 600 python bench.py --impl reference --steps 1 --warmup 0 2>/dev/null | tail -1 | cut -c1-600
} > gpurun_out/r2_probe27.txt 2>&1
tail -30 gpurun_out/r2_probe27.txt | cut -c1-2500
