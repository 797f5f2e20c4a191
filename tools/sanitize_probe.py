#!/usr/bin/env python
"""Small self-checking driver of the pass kernel in both launch shapes (written for
`compute-sanitizer --tool memcheck python tools/sanitize_probe.py`; the sanitizer is closed on this GPU pool,
so it is run plain -- gpurun_out/r1n_memcheck.txt records the refusal).
A 22-qubit state (2048 tiles of 11 qubits: warp-group shape) and a 16-qubit state (single-buffer shape) each
take a few passes with fused blocks, lone gates, diagonal runs, cx and a diagonal tail; the result is checked
against the numpy oracle so that the run also fails on wrong answers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import basicaer_oracle as orc
from qiskit_sdk_py_b200 import _lib
from qiskit_sdk_py_b200.engine import DeviceState
from qiskit_sdk_py_b200.lower import GateOp

rng = np.random.RandomState(11)


def unitary(k):
    q, r = np.linalg.qr(rng.standard_normal((2 ** k, 2 ** k)) + 1j * rng.standard_normal((2 ** k, 2 ** k)))
    return q * (np.diag(r) / np.abs(np.diag(r)))


def oracle_apply(vec, op, n):
    if op.kind in (_lib.GATE_DIAG1, _lib.GATE_DIAG2):
        mat = np.diag(op.matrix)
    elif op.kind == _lib.GATE_CX:
        mat = orc.cx_gate_matrix()
    else:
        mat = op.matrix
    return orc.apply_unitary(vec.reshape([2] * n), mat, list(op.qubits)).reshape(-1)


for n, want_threads in ((22, 768), (16, 256)):
    vec = rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)
    vec /= np.linalg.norm(vec)
    st = DeviceState(n)
    st.upload(vec)
    ref = vec.copy()
    for rep in range(2):
        tile = [0, 1, 2, 3] + sorted(int(q) for q in rng.choice(np.arange(4, n), 7, replace=False))
        t = tile
        ops = [GateOp(_lib.GATE_DENSE2, [t[4], t[9]], unitary(2)), GateOp(_lib.GATE_DENSE2, [t[9], t[1]], unitary(2)),
               GateOp(_lib.GATE_DENSE2, [t[6], t[3]], unitary(2)), GateOp(_lib.GATE_CX, [t[5], t[10]], None),
               GateOp(_lib.GATE_DIAG2, [t[7], t[8]], np.exp(1j * rng.uniform(0, 6, 4))),
               GateOp(_lib.GATE_DIAG1, [t[2]], np.exp(1j * rng.uniform(0, 6, 2))),
               GateOp(_lib.GATE_DIAG2, [t[0], t[8]], np.exp(1j * rng.uniform(0, 6, 4))),
               GateOp(_lib.GATE_DENSE1, [t[10]], unitary(1))]
        outside = [q for q in range(n) if q not in tile]
        tail = [GateOp(_lib.GATE_DIAG2, [outside[i % len(outside)], tile[i % 11]], np.exp(1j * rng.uniform(0, 6, 4)))
                for i in range(9)] if rep == 0 else None
        assert DeviceState.pass_info(n, tile, ops)["threads"] == want_threads
        st.apply_pass(tile, ops, tail=tail)
        for op in ops + (tail or []):
            ref = oracle_apply(ref, op, n)
    err = np.max(np.abs(st.download() - ref))
    print("n=%d threads=%d max-abs err %.2e" % (n, want_threads, err))
    assert err < 1e-12
print("sanitize probe ok")
