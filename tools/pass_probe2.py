#!/usr/bin/env python
"""Breakdown experiment: run the probe under QSV_DEBUG_FLAGS (1 no stores, 2 no loads, 4 no gates, 8 no pair
fusion).  Usage: pass_probe2.py [n_qubits] [comma-separated flag values]."""
import os, subprocess, sys
n = sys.argv[1] if len(sys.argv) > 1 else "30"
code = r'''
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from qiskit_sdk_py_b200 import _lib
from qiskit_sdk_py_b200.engine import DeviceState
from qiskit_sdk_py_b200.lower import GateOp
n = int(sys.argv[1]); st = DeviceState(n); st.init_basis0(); rng = np.random.RandomState(0)
def ru(k):
    z = rng.standard_normal((2**k, 2**k)) + 1j * rng.standard_normal((2**k, 2**k)); q, r = np.linalg.qr(z); return q * (np.diag(r) / np.abs(np.diag(r)))
def timeit(tile, ops, reps=5):
    for _ in range(2): st.apply_pass(tile, ops)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): st.apply_pass(tile, ops)
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
tile = list(range(4)) + list(range(n - 7, n))
pairs = [(0, 4), (5, 6), (7, 8), (9, 10), (1, 5), (2, 7), (3, 9), (4, 10), (6, 8), (0, 9)]
out = []
for k in (0, 2, 4, 8):
    ops = [GateOp(_lib.GATE_DENSE2, [tile[p[0]], tile[p[1]]], ru(2)) for p in pairs[:k]]
    out.append("%d gates %.3f ms" % (k, timeit(tile, ops)))
print(os.environ.get("QSV_DEBUG_FLAGS", "0"), " | ".join(out))
'''
for flags in (sys.argv[2].split(",") if len(sys.argv) > 2 else ("0", "1", "2", "3")):
    env = dict(os.environ, QSV_DEBUG_FLAGS=flags)
    subprocess.run([sys.executable, "-c", code, n], env=env, check=True)
