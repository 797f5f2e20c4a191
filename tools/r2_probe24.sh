#!/bin/bash
# round 2, GPU call 24 (2 GPUs): do pairwise swaps (k = 1) pay to overlap now that the exchange kernel gets 32 SMs and the
# gates an exchange does not need are carried over?  QuantumVolume(34), 33 local qubits
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  echo "== $name: $*"
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus 2 --steps 2 --warmup 1 --no-families --no-e2e --no-parity > gpurun_out/r2o_bench_2gpu_$name.json 2> gpurun_out/r2o_bench_2gpu_$name.err
  echo "rc=$?"; tail -2 gpurun_out/r2o_bench_2gpu_$name.err | cut -c1-300; python - <<PY
import json
d=json.load(open('gpurun_out/r2o_bench_2gpu_$name.json'))
print({k:d[k] for k in ("value","ms_per_step","schedule")})
PY
}
{
run default A=1
run ovl_k1_x32 QSV_DIST_MIN_OVL_SWAPS=1 QSV_DIST_XSMS=32
run ovl_k1_x48 QSV_DIST_MIN_OVL_SWAPS=1 QSV_DIST_XSMS=48
run ovl_k1_x32_g18 QSV_DIST_MIN_OVL_SWAPS=1 QSV_DIST_XSMS=32 QSV_DIST_OVERLAP_GATES=18
} > gpurun_out/r2_probe24.txt 2>&1
tail -40 gpurun_out/r2_probe24.txt | cut -c1-1500
