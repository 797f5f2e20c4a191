#!/usr/bin/env python
"""Time single passes of the tile kernel: copy-only, and k dense gates (marginal cost per gate)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from qiskit_sdk_py_b200 import _lib
from qiskit_sdk_py_b200.engine import DeviceState
from qiskit_sdk_py_b200.lower import GateOp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
st = DeviceState(n)
st.init_basis0()
rng = np.random.RandomState(0)

def ru(k):
    z = rng.standard_normal((2**k, 2**k)) + 1j * rng.standard_normal((2**k, 2**k))
    q, r = np.linalg.qr(z)
    return q * (np.diag(r) / np.abs(np.diag(r)))

def timeit(tile, ops, reps=5):
    for _ in range(2):
        st.apply_pass(tile, ops)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        st.apply_pass(tile, ops)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

tile = list(range(12))
gb = 32.0 * 2**n / 1e9
print("n=%d  one pass = %.2f GB" % (n, gb))
t0 = timeit(tile, [])
print("copy only (0 gates): %.3f ms  %.0f GB/s" % (t0, gb / t0 * 1e3))
pairs = [(3, 4), (5, 6), (7, 8), (9, 10), (11, 0), (1, 2), (3, 5), (4, 6), (7, 9), (8, 10), (11, 1), (0, 2)]
for k in (1, 2, 3, 4, 6, 8, 12):
    ops = [GateOp(_lib.GATE_DENSE2, list(p), ru(2)) for p in pairs[:k]]
    t = timeit(tile, ops)
    print("%2d dense2 gates: %.3f ms  %.0f GB/s  marginal %.3f ms/gate" % (k, t, gb / t * 1e3, (t - t0) / k))
# same-pair chain (single stage, no re-partition)
for k in (4, 8):
    ops = [GateOp(_lib.GATE_DENSE2, [3, 4], ru(2)) for _ in range(k)]
    t = timeit(tile, ops)
    print("%2d dense2 gates on one pair: %.3f ms  marginal %.3f ms/gate" % (k, t, (t - t0) / k))
ops = [GateOp(_lib.GATE_DIAG1, [q], np.exp(1j * rng.uniform(0, 6, 2))) for q in range(8)]
t = timeit(tile, ops)
print(" 8 diag1 gates: %.3f ms  marginal %.3f ms/gate" % (t, (t - t0) / 8))
ops = [GateOp(_lib.GATE_CX, list(p), None) for p in pairs[:8]]
t = timeit(tile, ops)
print(" 8 cx gates: %.3f ms  marginal %.3f ms/gate" % (t, (t - t0) / 8))
# high-qubit tile (strided 128-byte runs)
tile_hi = [0, 1, 2] + list(range(n - 9, n))
t = timeit(tile_hi, [])
print("copy only, tile = {0,1,2}+top 9: %.3f ms  %.0f GB/s" % (t, gb / t * 1e3))
for low in (3, 4, 5, 6):
    for b in (10, 11, 12):
        tl = list(range(low)) + list(range(n - (b - low), n))
        t = timeit(tl, [])
        print("copy only, b=%d tile = low %d + top %d: %.3f ms  %.0f GB/s" % (b, low, b - low, t, gb / t * 1e3))
for b in (10, 11, 12):
    t = timeit(list(range(b)), [])
    print("copy only, b=%d contiguous: %.3f ms  %.0f GB/s" % (b, t, gb / t * 1e3))
