#!/bin/bash
# round 2, GPU call 14 (4 GPUs): carry-over of the gates an exchange does not need (QSV_DIST_CARRY), SMs left to the
# exchange kernel under the sub-block passes (QSV_DIST_XSMS), overlap on/off -- QuantumVolume(35), 33 local qubits
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  echo "== $name: $*"
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus 4 --steps 2 --warmup 1 --no-families --no-e2e --no-parity > gpurun_out/r2h_bench_4gpu_$name.json 2> gpurun_out/r2h_bench_4gpu_$name.err
  echo "rc=$?"; tail -2 gpurun_out/r2h_bench_4gpu_$name.err | cut -c1-300; python - <<PY
import json
d=json.load(open('gpurun_out/r2h_bench_4gpu_$name.json'))
print({k:d[k] for k in ("value","ms_per_step","schedule")})
PY
}
{
run carry1_x16 QSV_DIST_CARRY=1 QSV_DIST_XSMS=16
run carry1_x32 QSV_DIST_CARRY=1 QSV_DIST_XSMS=32
run carry0_x16 QSV_DIST_CARRY=0 QSV_DIST_XSMS=16
run carry1_ovl0 QSV_DIST_CARRY=1 QSV_DIST_OVERLAP=0
} > gpurun_out/r2_probe14.txt 2>&1
tail -40 gpurun_out/r2_probe14.txt | cut -c1-1500
