#!/usr/bin/env python
"""Cost of a diagonal pass tail: one pass with 0 or 4 dense gates and N tail gates.  python tools/tail_probe.py [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from qiskit_sdk_py_b200 import _lib
from qiskit_sdk_py_b200.engine import DeviceState
from qiskit_sdk_py_b200.lower import GateOp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
st = DeviceState(n); st.init_basis0(); rng = np.random.RandomState(0)
def ru(k):
    z = rng.standard_normal((2**k, 2**k)) + 1j * rng.standard_normal((2**k, 2**k)); q, r = np.linalg.qr(z); return q * (np.diag(r) / np.abs(np.diag(r)))
def timeit(tile, ops, tail, reps=5):
    for _ in range(2): st.apply_pass(tile, ops, tail=tail)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): st.apply_pass(tile, ops, tail=tail)
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
tile = list(range(4)) + list(range(n - 7, n))
outside = [q for q in range(n) if q not in tile]
dense = [GateOp(_lib.GATE_DENSE2, [tile[a], tile[b]], ru(2)) for a, b in ((4, 5), (6, 7), (8, 9), (10, 0))]
for n_dense in (0, 4):
    row = []
    for n_tail in (0, 1, 8, 64, 190):
        tail = [GateOp(_lib.GATE_DIAG2, [tile[4 + i % 7], outside[i % len(outside)]], np.exp(1j * rng.uniform(0, 6, 4))) for i in range(n_tail)]
        row.append("tail %3d: %6.3f ms" % (n_tail, timeit(tile, dense[:n_dense], tail)))
    print("%d dense gates | " % n_dense + " | ".join(row))
