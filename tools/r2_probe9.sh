#!/bin/bash
# round 2, GPU call 9 (4 GPUs): overlapped exchange A/B at 33 local qubits (QuantumVolume(35))
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for cfg in "1 30" "0 30" "1 18"; do set -- $cfg; ovl=$1; tg=$2
echo "== bench --gpus 4 --qubits 33, overlap=$ovl target gates $tg"
QSV_DIST_OVERLAP=$ovl QSV_DIST_OVERLAP_GATES=$tg timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2957$ovl bench.py --gpus 4 --steps 2 --warmup 1 --no-families --no-e2e --no-parity > gpurun_out/r2d_bench_4gpu_ovl${ovl}_$tg.json 2> gpurun_out/r2d_bench_4gpu_ovl${ovl}_$tg.err
echo "rc=$?"; tail -2 gpurun_out/r2d_bench_4gpu_ovl${ovl}_$tg.err; python - <<PY
import json
d=json.load(open('gpurun_out/r2d_bench_4gpu_ovl${ovl}_$tg.json'))
print({k:d[k] for k in ("value","ms_per_step","schedule")}, d["check"])
PY
done
} > gpurun_out/r2_probe9.txt 2>&1
tail -30 gpurun_out/r2_probe9.txt | cut -c1-900
