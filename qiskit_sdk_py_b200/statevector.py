"""Device-resident counterpart of ``qiskit.quantum_info.Statevector`` for the calls that sit next to the hot path
(SURVEY 8(f) rank 4): opflow / VQE-style callers evaluate many Pauli expectation values on one final state, and with
the reference every one of them walks a host ndarray (``Statevector._expectation_value_pauli``,
quantum_info/states/statevector.py:377-408, Cython loops exp_value.pyx:37-94) after a 2^n-amplitude copy out of the
simulator.  ``DeviceStatevector`` keeps the amplitudes in HBM -- it wraps the engine state a simulator run leaves
behind, or builds one from instructions -- and answers

    expectation_value(pauli [, qargs])   Pauli label / qiskit Pauli / list of (label, coefficient) -> qsv_expval_pauli
    probabilities([qargs])               marginal probabilities in the reference's order        -> qsv_marginal
    sample_counts(shots, seed) / sample_memory                                                  -> exact-order sampler
    data / to_numpy()                    one D2H copy, only when asked for

with the reference's conventions (qubit 0 = rightmost label character, little-endian bit strings).  No CPU path:
constructing one needs the CUDA library and a GPU.
"""
import numpy as np

from . import lower
from .engine import DeviceState


def _pauli_label(pauli, num_qubits, qargs=None):
    """(full-width label, coefficient) for a label string or a qiskit ``Pauli`` (duck-typed: ``to_label()``; a leading
    sign / phase of the label such as '-iXZ' becomes the coefficient)."""
    label = pauli if isinstance(pauli, str) else pauli.to_label()
    coeff = 1.0 + 0.0j
    while label and label[0] in "+-i":
        coeff *= {"+": 1, "-": -1, "i": 1j}[label[0]]
        label = label[1:]
    label = label.upper()
    if any(c not in "IXYZ" for c in label):
        raise ValueError("not a Pauli label: %r" % (pauli,))
    if qargs is None:
        if len(label) != num_qubits:
            raise ValueError("Pauli on %d qubits, state has %d" % (len(label), num_qubits))
        return label, coeff
    qargs = list(qargs)
    if len(qargs) != len(label) or len(set(qargs)) != len(qargs):
        raise ValueError("qargs must name one distinct qubit per Pauli factor")
    full = ["I"] * num_qubits
    for pos, q in enumerate(qargs):  # label character for qargs[pos] is the pos-th from the RIGHT
        full[num_qubits - 1 - q] = label[len(label) - 1 - pos]
    return "".join(full), coeff


class DeviceStatevector:
    """2^n complex128 amplitudes resident on the GPU, with the read-only queries of ``Statevector``."""

    def __init__(self, state):
        """``state``: an ``engine.DeviceState`` (or any object with the same query methods, e.g. a sharded state)."""
        self._state = state

    # -- construction ---------------------------------------------------------------------
    @classmethod
    def from_amplitudes(cls, amplitudes):
        amps = np.ascontiguousarray(amplitudes, dtype=np.complex128).reshape(-1)
        n = int(round(np.log2(amps.size)))
        if amps.size != 1 << n:
            raise ValueError("length of the amplitude vector is not a power of two")
        st = DeviceState(n)
        st.upload(amps)
        return cls(st)

    @classmethod
    def from_instructions(cls, num_qubits, instructions, state_factory=None):
        """|0..0> evolved by Qobj-style gate instructions (dicts with name / qubits / params; basis gates and
        ``unitary``), entirely on the device."""
        st = (state_factory or DeviceState)(num_qubits)
        st.init_basis0()
        ops = [o for o in (lower.lower_gate(i["name"], i["qubits"], i.get("params")) for i in instructions) if o is not None]
        st.apply_ops(ops)
        return cls(st)

    def evolve(self, instructions):
        """Apply further gate instructions in place; returns self."""
        ops = [o for o in (lower.lower_gate(i["name"], i["qubits"], i.get("params")) for i in instructions) if o is not None]
        self._state.apply_ops(ops)
        return self

    # -- properties -----------------------------------------------------------------------
    @property
    def num_qubits(self):
        return self._state.n

    @property
    def dim(self):
        return 1 << self._state.n

    @property
    def device_state(self):
        return self._state

    def to_numpy(self):
        """Host copy (one D2H transfer of 16 * 2^n bytes)."""
        return self._state.download()

    data = property(to_numpy)

    def __array__(self, dtype=None, copy=None):
        arr = self.to_numpy()
        return arr if dtype is None else arr.astype(dtype)

    def __len__(self):
        return self.dim

    def norm(self):
        return float(np.sqrt(self._state.norm2()))

    # -- queries --------------------------------------------------------------------------
    def expectation_value(self, oper, qargs=None):
        """<psi|O|psi> for a Pauli (label string or qiskit ``Pauli``), or a linear combination given as a list of
        (pauli, coefficient) pairs / an object with ``to_list()`` (``SparsePauliOp``).  Mirrors
        ``Statevector.expectation_value`` for Pauli operators (statevector.py:377-421): one streaming reduction over
        the device state per Pauli term; nothing but the scalar leaves the GPU."""
        if hasattr(oper, "to_list") and not isinstance(oper, str):
            oper = oper.to_list()
        if isinstance(oper, (list, tuple)) and oper and isinstance(oper[0], (list, tuple)):
            total = 0.0 + 0.0j
            for pauli, coeff in oper:
                total += complex(coeff) * self.expectation_value(pauli, qargs)
            return total
        label, coeff = _pauli_label(oper, self.num_qubits, qargs)
        return coeff * self._state.expval_pauli(label)

    def probabilities(self, qargs=None):
        """Marginal probabilities of ``qargs`` (all qubits if None) as a float64 ndarray of 2^m entries; bit pos of
        the index <-> qargs[pos], as ``Statevector.probabilities``."""
        n = self.num_qubits
        qargs = list(range(n)) if qargs is None else [int(q) for q in qargs]
        order = sorted(range(len(qargs)), key=lambda i: qargs[i])
        measured_sorted = [qargs[i] for i in order]
        probs = self._state.marginal(measured_sorted).cpu().numpy()
        if measured_sorted == qargs:
            return probs
        # permute the index bits: bit `pos_sorted` of the device result is qubit measured_sorted[pos_sorted]
        m = len(qargs)
        idx = np.arange(1 << m)
        src = np.zeros_like(idx)
        for pos, i in enumerate(order):  # sorted position pos holds the qubit asked for at position i
            src |= ((idx >> i) & 1) << pos
        return probs[src]

    def sample_memory(self, shots, qargs=None, seed=None):
        """``shots`` bit strings (qubit qargs[0] rightmost), drawn like the reference's sampler: numpy
        ``RandomState(seed).random_sample(shots)`` looked up in the exact-order CDF on the device."""
        n = self.num_qubits
        qargs = list(range(n)) if qargs is None else [int(q) for q in qargs]
        measured_sorted = sorted(set(qargs))
        uniforms = np.random.RandomState(seed).random_sample(int(shots))
        samples, _total = self._state.sample(measured_sorted, uniforms, exact_order=True)
        pos = {q: i for i, q in enumerate(measured_sorted)}
        out = []
        for s in samples.tolist():
            out.append("".join(str((s >> pos[q]) & 1) for q in reversed(qargs)))
        return out

    def sample_counts(self, shots, qargs=None, seed=None):
        counts = {}
        for key in self.sample_memory(shots, qargs, seed):
            counts[key] = counts.get(key, 0) + 1
        return counts
