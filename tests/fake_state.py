"""TEST DOUBLE for engine.DeviceState, backed by the numpy oracle.

Lets the CPU test-suite exercise the *host* logic of the product (Qobj lowering, fusion planner,
classical control flow, RNG consumption, result layout) without a GPU.  It lives under tests/ and is
never importable from the package: the product's default state factory is the CUDA engine.
"""
import numpy as np

from oracle import basicaer_oracle as orc
from qiskit_sdk_py_b200 import _lib
from qiskit_sdk_py_b200.planner import merge_ops, plan_passes


class FakeState:
    def __init__(self, n_qubits, tile_qubits=6):
        self.n = n_qubits
        self.tile_qubits = tile_qubits
        self.vec = np.zeros(2 ** n_qubits, dtype=complex)
        self.passes_launched = 0
        self.gates_applied = 0
        self.pass_log = []

    def init_basis0(self, phase=1.0 + 0.0j):
        self.vec = np.zeros(2 ** self.n, dtype=complex)
        self.vec[0] = phase

    def init_identity(self, n_half, phase=1.0 + 0.0j):
        self.vec = (np.eye(2 ** n_half, dtype=complex) * phase).reshape(-1)

    def upload(self, amplitudes, phase=1.0 + 0.0j):
        self.vec = np.array(amplitudes, dtype=complex).reshape(-1) * phase

    def _apply(self, op):
        t = self.vec.reshape([2] * self.n) if self.n else self.vec
        if op.kind in (_lib.GATE_DIAG1, _lib.GATE_DIAG2):
            mat = np.diag(op.matrix)
        elif op.kind == _lib.GATE_CX:
            mat = orc.cx_gate_matrix()
        else:
            mat = op.matrix
        self.vec = orc.apply_unitary(t, mat, list(op.qubits)).reshape(-1)

    def apply_ops(self, ops):
        run = []
        for op in ops:
            if op.kind == "densek":
                self._flush(run)
                run = []
                self._apply(op)
                self.passes_launched += 1
                self.gates_applied += 1
            else:
                run.append(op)
        self._flush(run)

    def _flush(self, run):
        if not run:
            return
        for p in plan_passes(merge_ops(run), self.n, self.tile_qubits, tails=True):
            self.apply_pass(p.tile_qubits, p.ops, tail=p.tail)
            self.pass_log.append(p)

    def apply_pass(self, tile_qubits, ops, dest=None, tail=None):
        assert dest is None
        assert len(tile_qubits) == min(self.n, self.tile_qubits) and len(set(tile_qubits)) == len(tile_qubits)
        for op in ops:
            assert set(op.qubits) <= set(tile_qubits)
            self._apply(op)
        tail = tail or []
        assert len(tail) <= _lib.MAX_TAIL_GATES
        for op in tail:  # the contract of qsv_apply_pass_tail
            assert op.kind in (_lib.GATE_DIAG1, _lib.GATE_DIAG2)
            assert len(op.qubits) == 1 or not set(op.qubits) <= set(tile_qubits)
            self._apply(op)
        self.passes_launched += 1
        self.gates_applied += len(ops) + len(tail)

    def plan_ops(self, ops, avoid=()):
        return plan_passes(merge_ops(list(ops)), self.n, self.tile_qubits, tails=True, avoid=avoid)

    def apply_pass_sub(self, tile_qubits, ops, tail, fixed_qubits, pattern, max_sms=0):
        """The contract of qsv_apply_pass_sub: only the amplitudes whose fixed qubits read ``pattern`` change; the
        fixed qubits are neither tile qubits nor gate qubits."""
        assert not set(fixed_qubits) & set(tile_qubits)
        assert len(tile_qubits) == min(self.n, self.tile_qubits)
        before = self.vec.copy()
        for op in list(ops) + list(tail or []):
            assert not set(op.qubits) & set(fixed_qubits)
            self._apply(op)
        idx = np.arange(self.vec.size)
        keep = np.ones(self.vec.size, dtype=bool)
        for i, q in enumerate(fixed_qubits):
            keep &= ((idx >> q) & 1) == ((pattern >> i) & 1)
        self.vec = np.where(keep, self.vec, before)

    def prob_qubit(self, qubit):
        p = np.abs(self.vec) ** 2
        idx = np.arange(p.size)
        mask = ((idx >> qubit) & 1).astype(bool)
        return float(p[~mask].sum()), float(p[mask].sum())

    def collapse(self, qubit, outcome, prob, is_reset=False):
        s = 1 / np.sqrt(prob)
        if is_reset and outcome == 1:
            mat = [[0, s], [0, 0]]
        elif outcome == 0:
            mat = [[s, 0], [0, 0]]
        else:
            mat = [[0, 0], [0, s]]
        self.vec = orc.apply_unitary(self.vec.reshape([2] * self.n), mat, [qubit]).reshape(-1)

    # shot-batched mid-circuit operations (engine.DeviceState.collapse_batch / apply_masked)
    def collapse_batch(self, n_shot_qubits, qubit, outcomes, scales, is_reset=False):
        n = self.n - n_shot_qubits
        vec = self.vec.reshape(1 << n_shot_qubits, 1 << n).copy()
        idx = np.arange(1 << n)
        bit = (idx >> qubit) & 1
        for s in range(1 << n_shot_qubits):
            out, f = int(outcomes[s]), float(scales[s])
            row = vec[s]
            if is_reset and out == 1:
                new = np.zeros_like(row)
                new[idx[bit == 0]] = row[idx[bit == 0] | (1 << qubit)] * f
            else:
                new = np.where(bit == out, row * f, 0.0)
            vec[s] = new
        self.vec = vec.reshape(-1)
        self.passes_launched += 1

    def apply_masked(self, n_shot_qubits, op, mask):
        n = self.n - n_shot_qubits
        before = self.vec.copy()
        self._apply(op)
        keep = np.repeat(np.asarray(mask, dtype=bool), 1 << n)
        self.vec = np.where(keep, self.vec, before)
        self.passes_launched += 1
        self.gates_applied += 1

    def scale(self, factor):
        self.vec = self.vec * complex(factor)

    def expval_pauli(self, pauli):
        return orc.expval_pauli(self.vec, pauli)

    def norm2(self):
        return float(np.sum(np.abs(self.vec) ** 2))

    def comm_device(self):
        return "cpu"

    def _half_index(self, local_qubit, bit_value, begin, count):
        h = np.arange(begin, begin + count, dtype=np.int64)
        low = h & ((1 << local_qubit) - 1)
        return ((h >> local_qubit) << (local_qubit + 1)) | (bit_value << local_qubit) | low

    def pack_half(self, local_qubit, bit_value, buf, begin, count):
        import torch

        idx = self._half_index(local_qubit, bit_value, begin, count)
        buf[: 2 * count] = torch.from_numpy(np.ascontiguousarray(self.vec[idx]).view(np.float64).copy())

    def unpack_half(self, local_qubit, bit_value, buf, begin, count):
        idx = self._half_index(local_qubit, bit_value, begin, count)
        self.vec[idx] = buf[: 2 * count].numpy().copy().view(np.complex128)

    def _sub_index(self, local_qubits, pattern, begin, count):
        h = np.arange(begin, begin + count, dtype=np.int64)
        order = sorted(range(len(local_qubits)), key=lambda i: local_qubits[i])
        value = 0
        for i in order:
            q = local_qubits[i]
            h = ((h >> q) << (q + 1)) | (h & ((1 << q) - 1))
            value |= ((pattern >> i) & 1) << q
        return h | value

    def pack_bits(self, local_qubits, pattern, buf, begin, count):
        import torch

        idx = self._sub_index(local_qubits, pattern, begin, count)
        buf[: 2 * count] = torch.from_numpy(np.ascontiguousarray(self.vec[idx]).view(np.float64).copy())

    def unpack_bits(self, local_qubits, pattern, buf, begin, count):
        idx = self._sub_index(local_qubits, pattern, begin, count)
        self.vec[idx] = buf[: 2 * count].numpy().copy().view(np.complex128)

    def _marginal_np(self, measured_sorted):
        n = self.n
        axis = list(range(n))
        for q in reversed(measured_sorted):
            axis.remove(n - 1 - q)
        return np.reshape(np.sum(np.abs(self.vec.reshape([2] * n)) ** 2, axis=tuple(axis)),
                          2 ** len(measured_sorted))

    def marginal(self, measured_sorted):
        import torch

        return torch.from_numpy(np.ascontiguousarray(self._marginal_np(list(measured_sorted)), dtype=np.float64))

    def scatter_bits(self, src, dest_bits, base, log2_dst):
        import torch

        s = np.arange(1 << len(dest_bits), dtype=np.int64)
        dest = np.full_like(s, int(base))
        for i, b in enumerate(dest_bits):
            dest |= ((s >> i) & 1) << b
        out = np.zeros(1 << log2_dst)
        out[dest] = src.numpy()
        return torch.from_numpy(out)

    def sample(self, measured_sorted, uniforms, exact_order=True, check_norm=True):
        return self._sample_np(self._marginal_np(measured_sorted), uniforms)

    def sample_probs(self, probs, log2_bins, uniforms, exact_order=True, check_norm=True):
        p = probs.numpy().copy()
        assert p.size == 1 << log2_bins
        return self._sample_np(p, uniforms)

    @staticmethod
    def _sample_np(p, uniforms):
        cdf = p.cumsum()
        total = cdf[-1]
        cdf /= total
        return cdf.searchsorted(uniforms, side="right").astype(np.int64), float(total)

    # the exact-order sampler phase by phase (engine.DeviceState.cdf_*): numpy itself, continued across shards
    def cdf_prepare(self, shots):
        self._p = np.abs(self.vec) ** 2
        return float(self._p.sum())

    def cdf_classify(self, approx_offset, shots):
        pass

    def cdf_chain(self, exact_start, shots):
        # cumsum([start, p0, p1, ...]) is the one running sum numpy would compute over the concatenated shards
        self._cdf = np.cumsum(np.concatenate(([exact_start], self._p)))[1:]
        return float(self._cdf[-1])

    def cdf_search(self, exact_start, total, uniforms):
        uniforms = np.asarray(uniforms, dtype=np.float64)
        idx = (self._cdf / total).searchsorted(uniforms, side="right").astype(np.int64)
        idx[idx >= self._cdf.size] = -2
        idx[exact_start / total > uniforms] = -1
        return idx

    def chop(self, thr):
        self.vec[abs(self.vec) < thr] = 0.0

    def download(self):
        return self.vec.copy()

    def snapshot(self):
        return self.vec.copy()

    def restore(self, snap):
        self.vec = snap.copy()
