#!/bin/bash
cd "$(dirname "$0")/.."
DBG=$PWD/qiskit_sdk_py_b200/libqsv_dbg.so
mkdir -p gpurun_out
{
echo "== dmma_smem"; timeout 120 tools/bin/dmma_smem
echo "== phase + per-op profile"; QSV_LIB_PATH=$DBG timeout 300 python tools/prof_dump.py 30 12
} > gpurun_out/r2_probe3.txt 2>&1
tail -50 gpurun_out/r2_probe3.txt
