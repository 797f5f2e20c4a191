#!/bin/bash
# round 2, GPU call 1: correctness of the producer-warp kernel, microbenchmarks, A/B against the round-1 shape
cd "$(dirname "$0")/.."
DBG=$PWD/qiskit_sdk_py_b200/libqsv_dbg.so
mkdir -p gpurun_out
{
echo "== pytest gpu (production lib)"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== dmma_lat"; timeout 120 tools/bin/dmma_lat
echo "== tma_issue"; timeout 120 tools/bin/tma_issue
echo "== QV-30 variants (dbg lib)"; QSV_LIB_PATH=$DBG bash tools/b30_variants.sh "producer A=1" "round1_shape QSV_DEBUG_FLAGS=512" "producer_again A=1"
echo "== pass breakdown"; QSV_LIB_PATH=$DBG timeout 300 python tools/pass_probe2.py 30 0,3,1,2,512,515
echo "== phase profile"; QSV_LIB_PATH=$DBG timeout 300 python tools/prof_dump.py 30 14
} > gpurun_out/r2_probe1.txt 2>&1
tail -60 gpurun_out/r2_probe1.txt
