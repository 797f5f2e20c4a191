#!/bin/bash
# round 2, GPU call 2: software-pipelined DMMA paths (A/B against the previous build on the same box), new sampler tests
cd "$(dirname "$0")/.."
DBG=$PWD/qiskit_sdk_py_b200/libqsv_dbg.so
PREV=$PWD/qiskit_sdk_py_b200/libqsv_prev.so
mkdir -p gpurun_out
{
echo "== pytest gpu kernels (production lib)"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -8
echo "== QV-30 variants"; 
QSV_LIB_PATH=$PREV bash tools/b30_variants.sh "prev_producer A=1"
QSV_LIB_PATH=$DBG bash tools/b30_variants.sh "pipelined A=1"
QSV_LIB_PATH=$PREV bash tools/b30_variants.sh "prev_producer_again A=1"
QSV_LIB_PATH=$DBG bash tools/b30_variants.sh "pipelined_again A=1" "pipelined_round1_shape QSV_DEBUG_FLAGS=512"
echo "== phase profile (pipelined)"; QSV_LIB_PATH=$DBG timeout 300 python tools/prof_dump.py 30 14
echo "== pass breakdown"; QSV_LIB_PATH=$DBG timeout 300 python tools/pass_probe2.py 30 0,3
} > gpurun_out/r2_probe2.txt 2>&1
tail -40 gpurun_out/r2_probe2.txt
