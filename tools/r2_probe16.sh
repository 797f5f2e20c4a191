#!/bin/bash
# round 2, GPU call 16 (4 GPUs): fused exchange as one kernel vs partner by partner (QSV_DIST_STEPWISE), without overlap;
# 48 SMs for the exchange kernel under the sub-block passes -- QuantumVolume(35), 33 local qubits
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  echo "== $name: $*"
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus 4 --steps 2 --warmup 1 --no-families --no-e2e --no-parity > gpurun_out/r2j_bench_4gpu_$name.json 2> gpurun_out/r2j_bench_4gpu_$name.err
  echo "rc=$?"; tail -2 gpurun_out/r2j_bench_4gpu_$name.err | cut -c1-300; python - <<PY
import json
d=json.load(open('gpurun_out/r2j_bench_4gpu_$name.json'))
print({k:d[k] for k in ("value","ms_per_step","schedule")})
PY
}
{
run ovl0_step0 QSV_DIST_OVERLAP=0 QSV_DIST_STEPWISE=0
run ovl0_step1 QSV_DIST_OVERLAP=0 QSV_DIST_STEPWISE=1
run ovl1_x48_step1 QSV_DIST_XSMS=48 QSV_DIST_STEPWISE=1
run ovl1_x32_step1 QSV_DIST_XSMS=32 QSV_DIST_STEPWISE=1
} > gpurun_out/r2_probe16.txt 2>&1
tail -40 gpurun_out/r2_probe16.txt | cut -c1-1500
