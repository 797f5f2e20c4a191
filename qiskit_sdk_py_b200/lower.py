"""Lower Qobj instructions to device gate ops.

Host-side restatement of the tiny gate-matrix helpers of the reference
(qiskit/providers/basicaer/basicaertools.py:26-72) -- they stay on the host as the survey prescribes
(a3/a4: microseconds per gate, bit-identical matrices for free) -- plus the classification of each
matrix into the kernel's gate kinds (dense / diagonal / cx).
"""
import numpy as np

from . import _lib

# qiskit/providers/basicaer/basicaertools.py:26
SINGLE_QUBIT_GATES = ("U", "u1", "u2", "u3", "rz", "sx", "x")


class GateOp:
    """One gate for the device: ``kind`` is a _lib.GATE_* code or "densek" (k >= 3 qubits)."""

    __slots__ = ("kind", "qubits", "matrix")

    def __init__(self, kind, qubits, matrix):
        self.kind = kind
        self.qubits = tuple(int(q) for q in qubits)
        self.matrix = matrix  # complex128 ndarray: (2^k,2^k) dense, (2^k,) diagonal, None for cx

    @property
    def payload_doubles(self):
        # doubles of matrix payload inside a pass (a cx runs as its 4x4 permutation matrix on the tensor path)
        return {_lib.GATE_DENSE1: 32, _lib.GATE_DENSE2: 32, _lib.GATE_DIAG1: 4, _lib.GATE_DIAG2: 8,
                _lib.GATE_CX: 32}.get(self.kind, 0)

    def __repr__(self):
        return "GateOp(%r, %r)" % (self.kind, self.qubits)


def single_gate_matrix(name, params=None):
    """2x2 matrix of a basis gate; same closed forms (and the same numpy calls, so the same bits) as
    the ``__array__`` methods that basicaertools.py:29-63 reaches through ``Gate.to_matrix()``:
    u.py:98-110, u3.py:98-110, u2.py:85-96, u1.py:121-124, rz.py:103-108, sx.py:100-102, x.py:115-117."""
    params = [] if params is None else [float(p) for p in params]
    if name in ("U", "u3"):
        theta, phi, lam = params
        cos = np.cos(theta / 2)
        sin = np.sin(theta / 2)
        return np.array(
            [[cos, -np.exp(1j * lam) * sin], [np.exp(1j * phi) * sin, np.exp(1j * (phi + lam)) * cos]],
            dtype=complex,
        )
    if name == "u2":
        phi, lam = params
        isqrt2 = 1 / np.sqrt(2)
        return np.array(
            [[isqrt2, -np.exp(1j * lam) * isqrt2],
             [np.exp(1j * phi) * isqrt2, np.exp(1j * (phi + lam)) * isqrt2]],
            dtype=complex,
        )
    if name == "u1":
        (lam,) = params
        return np.array([[1, 0], [0, np.exp(1j * lam)]], dtype=complex)
    if name == "rz":
        (lam,) = params
        ilam2 = 0.5j * lam
        return np.array([[np.exp(-ilam2), 0], [0, np.exp(ilam2)]], dtype=complex)
    if name == "sx":
        return np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]], dtype=complex) / 2
    if name == "x":
        return np.array([[0, 1], [1, 0]], dtype=complex)
    raise ValueError("Gate is not a valid basis gate for this simulator: %s" % name)


def matrix_op(matrix, qubits):
    """Classify a k-qubit matrix acting on ``qubits`` (qubits[0] = LSB of the matrix index,
    qasm_simulator.py:145-161)."""
    k = len(qubits)
    mat = np.asarray(matrix, dtype=complex)
    if mat.shape != (2 ** k, 2 ** k):
        mat = mat.reshape(2 ** k, 2 ** k)
    if len(set(qubits)) != k:
        raise ValueError("duplicate qubits in gate: %r" % (qubits,))
    if k <= 2:
        off_diag = mat - np.diag(np.diag(mat))
        if not np.any(off_diag):
            return GateOp(_lib.GATE_DIAG1 if k == 1 else _lib.GATE_DIAG2, qubits, np.diag(mat).copy())
        return GateOp(_lib.GATE_DENSE1 if k == 1 else _lib.GATE_DENSE2, qubits, mat)
    return GateOp("densek", qubits, mat)


def lower_gate(name, qubits, params):
    """Qobj gate instruction -> GateOp, or None for the no-ops (id/u0/barrier;
    qasm_simulator.py:553-554,565-566)."""
    if name == "unitary":
        return matrix_op(params[0], qubits)
    if name in SINGLE_QUBIT_GATES:
        mat = single_gate_matrix(name, params)
        return matrix_op(mat, [qubits[0]])
    if name in ("CX", "cx"):
        if qubits[0] == qubits[1]:
            raise ValueError("duplicate qubits in cx")
        return GateOp(_lib.GATE_CX, [qubits[0], qubits[1]], None)
    if name in ("id", "u0", "barrier"):
        return None
    raise KeyError(name)


def shift_op(op, offset):
    """The same gate on qubits + offset (unitary simulator: gates act on the row bits n..2n-1)."""
    return GateOp(op.kind, [q + offset for q in op.qubits], op.matrix)
