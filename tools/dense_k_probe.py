#!/usr/bin/env python
"""Time the k-qubit dense gate (`unitary` instruction) for k = 2..8 at n qubits: ms per gate, HBM GB/s of the pass
(32 * 2^n bytes) and FP64 TF/s of the executed work (k = 3: 49 flop/amplitude; k >= 4: 3 real 2^k x 2^k products =
6 * 2^k flop/amplitude).  python tools/dense_k_probe.py [n_qubits]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from qiskit_sdk_py_b200 import lower  # noqa: E402
from qiskit_sdk_py_b200.engine import DeviceState  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.RandomState(0)
st = DeviceState(n)
st.init_basis0()
for k in range(2, 9):
    z = rng.standard_normal((2 ** k, 2 ** k)) + 1j * rng.standard_normal((2 ** k, 2 ** k))
    q, r = np.linalg.qr(z)
    mat = q * (np.diag(r) / np.abs(np.diag(r)))
    qubits = [int(x) for x in rng.choice(np.arange(4, n), k, replace=False)]
    op = lower.matrix_op(mat, qubits)
    for _ in range(2):
        st.apply_ops([op])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        st.apply_ops([op])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = {2: 32.0, 3: 49.0}.get(k, 6.0 * 2 ** k) * 2 ** n
    print("k=%d qubits %s: %.3f ms  %.0f GB/s  %.1f TF/s (executed FP64)%s" % (
        k, qubits, ms, 32.0 * 2 ** n / ms / 1e6, flops / ms / 1e9,
        "  [includes the host-side fragment set-up + H2D copy of the matrix]" if k >= 4 else ""))
print("norm2", st.norm2())
