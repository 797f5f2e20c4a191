"""Device-resident statevector: the qiskit-free engine that drives libqsv through ctypes.

PyTorch is used only as plumbing: it owns the device allocation (one float64 tensor of 2*2^n
entries = numpy complex128 layout) and the CUDA stream; every arithmetic operation on the state is
a hand-written sm_100a kernel of libqsv (include/qsv.h).  No CPU fallback: constructing a
``DeviceState`` without a CUDA device or without libqsv.so raises.
"""
import ctypes

import numpy as np

from . import _lib
from .lower import GateOp
from .planner import DEFAULT_TILE_QUBITS, eliminate_swaps, merge_ops, plan_passes

_p_int32 = ctypes.POINTER(ctypes.c_int32)


def _int32_array(values):
    arr = (ctypes.c_int32 * len(values))(*[int(v) for v in values])
    return arr


def _fill_gate(dst, op):
    """GateOp -> qsv_gate_t."""
    dst.kind = op.kind
    dst.q0 = op.qubits[0]
    dst.q1 = op.qubits[1] if len(op.qubits) > 1 else 0
    dst.reserved = 0
    if op.matrix is not None:
        flat = np.ascontiguousarray(op.matrix, dtype=np.complex128).reshape(-1).view(np.float64)
        ctypes.memmove(dst.m, flat.ctypes.data, flat.nbytes)


class DeviceState:
    """2^n complex128 amplitudes in HBM plus the small workspaces the kernels need."""

    def __init__(self, n_qubits, device=None, tile_qubits=DEFAULT_TILE_QUBITS):
        import torch

        self._torch = torch
        self.lib = _lib.load()  # raises QsvLibraryError if the CUDA library is not built
        if not torch.cuda.is_available():
            raise _lib.QsvLibraryError(
                "no CUDA device available: this backend runs only on a GPU (no CPU fallback)")
        self.n = int(n_qubits)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.tile_qubits = int(tile_qubits)
        with torch.cuda.device(self.device):
            self.state = torch.empty(2 << self.n, dtype=torch.float64, device=self.device)
            self._reduce_ws = torch.empty(self.lib.qsv_reduce_workspace_bytes() // 8 + 8,
                                          dtype=torch.float64, device=self.device)
        self._mat_scratch = None
        self._tail_scratch = None
        self._sample_ws = None
        self._marg_ws = None
        self._probs = None
        self.passes_launched = 0
        self.gates_applied = 0

    # -- plumbing --------------------------------------------------------------------
    @property
    def ptr(self):
        return ctypes.c_void_p(self.state.data_ptr())

    @property
    def stream(self):
        return ctypes.c_void_p(self._torch.cuda.current_stream(self.device).cuda_stream)

    def _ctx(self):
        return self._torch.cuda.device(self.device)

    def synchronize(self):
        self._torch.cuda.synchronize(self.device)

    # -- initialisation --------------------------------------------------------------
    def init_basis0(self, phase=1.0 + 0.0j):
        with self._ctx():
            _lib.check(self.lib.qsv_init_basis0(self.ptr, self.n, phase.real, phase.imag, self.stream),
                       "qsv_init_basis0")

    def init_identity(self, n_half, phase=1.0 + 0.0j):
        assert 2 * n_half == self.n
        with self._ctx():
            _lib.check(self.lib.qsv_init_identity(self.ptr, n_half, phase.real, phase.imag, self.stream),
                       "qsv_init_identity")

    def upload(self, amplitudes, phase=1.0 + 0.0j):
        host = np.ascontiguousarray(amplitudes, dtype=np.complex128).reshape(-1)
        if host.size != (1 << self.n):
            raise ValueError("upload: expected %d amplitudes, got %d" % (1 << self.n, host.size))
        with self._ctx():
            _lib.check(self.lib.qsv_upload(self.ptr, self.n, ctypes.c_void_p(host.ctypes.data), host.size,
                                           phase.real, phase.imag, self.stream), "qsv_upload")
            self.synchronize()  # host buffer may be released by the caller

    def scale(self, factor):
        factor = complex(factor)
        with self._ctx():
            _lib.check(self.lib.qsv_scale(self.ptr, self.n, factor.real, factor.imag, self.stream), "qsv_scale")

    # -- gates -----------------------------------------------------------------------
    def apply_pass(self, tile_qubits, ops, dest=None, tail=None):
        """One cache-blocked pass on PHYSICAL qubits; ``dest`` (optional) permutes the tile qubits on
        write-back (content of tile_qubits[i] lands on dest[i]); ``tail`` (optional) is a list of diagonal
        GateOps, on any qubits, applied after ``ops`` inside the same pass (qsv_apply_pass_tail)."""
        gates = (_lib.QsvGate * max(1, len(ops)))()
        for i, op in enumerate(ops):
            _fill_gate(gates[i], op)
        tile = _int32_array(tile_qubits)
        with self._ctx():
            if tail:
                tgates = (_lib.QsvGate * len(tail))()
                for i, op in enumerate(tail):
                    _fill_gate(tgates[i], op)
                if self._tail_scratch is None:
                    self._tail_scratch = self._torch.empty(int(self.lib.qsv_tail_scratch_bytes()) // 8,
                                                           dtype=self._torch.float64, device=self.device)
                _lib.check(self.lib.qsv_apply_pass_tail(self.ptr, self.n, tile, len(tile_qubits), gates, len(ops),
                                                        None if dest is None else _int32_array(dest), tgates,
                                                        len(tail), ctypes.c_void_p(self._tail_scratch.data_ptr()),
                                                        self.stream), "qsv_apply_pass_tail")
                self.gates_applied += len(tail)
            elif dest is None:
                _lib.check(self.lib.qsv_apply_pass(self.ptr, self.n, tile, len(tile_qubits), gates, len(ops),
                                                   self.stream), "qsv_apply_pass")
            else:
                _lib.check(self.lib.qsv_apply_pass_remap(self.ptr, self.n, tile, len(tile_qubits), gates,
                                                         len(ops), _int32_array(dest), self.stream),
                           "qsv_apply_pass_remap")
        self.passes_launched += 1
        self.gates_applied += len(ops)

    def apply_pass_sub(self, tile_qubits, ops, tail, fixed_qubits, pattern, max_sms=0):
        """One pass restricted to the sub-block whose qubits ``fixed_qubits[i]`` read bit i of ``pattern``, on at most
        ``max_sms`` SMs (qsv_apply_pass_sub)."""
        gates = (_lib.QsvGate * max(1, len(ops)))()
        for i, op in enumerate(ops):
            _fill_gate(gates[i], op)
        tail = tail or []
        tgates = (_lib.QsvGate * max(1, len(tail)))()
        for i, op in enumerate(tail):
            _fill_gate(tgates[i], op)
        with self._ctx():
            if tail and self._tail_scratch is None:
                self._tail_scratch = self._torch.empty(int(self.lib.qsv_tail_scratch_bytes()) // 8,
                                                       dtype=self._torch.float64, device=self.device)
            scratch = ctypes.c_void_p(self._tail_scratch.data_ptr()) if tail else ctypes.c_void_p(0)
            _lib.check(self.lib.qsv_apply_pass_sub(self.ptr, self.n, _int32_array(tile_qubits), len(tile_qubits), gates,
                                                   len(ops), tgates, len(tail), scratch, _int32_array(fixed_qubits),
                                                   len(fixed_qubits), int(pattern), int(max_sms), self.stream),
                       "qsv_apply_pass_sub")

    def plan_ops(self, ops, avoid=()):
        """The passes apply_ops would run for 1-/2-qubit ``ops`` (host only); tile qubits never taken from ``avoid``
        unless a gate needs them."""
        return plan_passes(merge_ops(list(ops)), self.n, self.tile_qubits, tails=True, avoid=avoid)

    @property
    def sm_count(self):
        return self._torch.cuda.get_device_properties(self.device).multi_processor_count

    @staticmethod
    def pass_info(n_qubits, tile_qubits, ops, dest=None):
        """Host-only: how qsv_apply_pass would lower this pass (no GPU needed).  Returns a dict with the
        kernel ops after gate fusion, stages, fused 3-qubit blocks, shared memory, threads, copy piece sizes."""
        lib = _lib.load()
        gates = (_lib.QsvGate * max(1, len(ops)))()
        for i, op in enumerate(ops):
            _fill_gate(gates[i], op)
        info = (ctypes.c_int32 * 8)()
        _lib.check(lib.qsv_pass_info(n_qubits, _int32_array(tile_qubits), len(tile_qubits), gates, len(ops),
                                     None if dest is None else _int32_array(dest), info), "qsv_pass_info")
        keys = ("ops", "stages", "fused3", "smem_bytes", "threads", "piece_log2_load", "piece_log2_store", "gates")
        return dict(zip(keys, list(info)))

    @staticmethod
    def pass_ops(n_qubits, tile_qubits, ops, dest=None):
        """Host-only: the kernel ops of the lowered pass, as dicts (kind, stage, qubits, conflict-free flags)."""
        lib = _lib.load()
        gates = (_lib.QsvGate * max(1, len(ops)))()
        for i, op in enumerate(ops):
            _fill_gate(gates[i], op)
        out = (ctypes.c_int32 * (4 * _lib.MAX_PASS_GATES))()
        rc = lib.qsv_pass_ops(n_qubits, _int32_array(tile_qubits), len(tile_qubits), gates, len(ops),
                              None if dest is None else _int32_array(dest), out, _lib.MAX_PASS_GATES)
        if rc < 0:
            _lib.check(rc, "qsv_pass_ops")
        return [{"kind": out[4 * o], "stage": out[4 * o + 1], "q0": out[4 * o + 2] & 255, "q1": out[4 * o + 2] >> 8,
                 "flags": out[4 * o + 3]} for o in range(rc)]

    def apply_dense(self, qubits, matrix):
        """One stand-alone k-qubit gate (the `unitary` instruction with k >= 3)."""
        k = len(qubits)
        mat = np.ascontiguousarray(matrix, dtype=np.complex128).reshape(2 ** k, 2 ** k)
        if k > _lib.MAX_DENSE_K:
            raise _lib.QsvLibraryError("unitary on %d qubits: more than %d qubits is not supported"
                                       % (k, _lib.MAX_DENSE_K))
        need = 4 * 4 ** k
        if k >= 4 and (self._mat_scratch is None or self._mat_scratch.numel() < need):
            self._mat_scratch = self._torch.empty(need, dtype=self._torch.float64, device=self.device)
        scratch = ctypes.c_void_p(self._mat_scratch.data_ptr()) if k >= 4 else ctypes.c_void_p(0)
        with self._ctx():
            _lib.check(self.lib.qsv_apply_matrix(self.ptr, self.n, _int32_array(qubits), k,
                                                 ctypes.c_void_p(mat.ctypes.data), scratch, self.stream),
                       "qsv_apply_matrix")
            if k >= 4:
                self.synchronize()  # the pageable host matrix was staged by the async copy; be safe
        self.passes_launched += 1
        self.gates_applied += 1

    def apply_ops(self, ops):
        """Apply a list of GateOps in program order: runs of 1-/2-qubit gates are fused into
        cache-blocked passes, k>=3 dense gates are stand-alone passes."""
        run = []
        for op in ops:
            if op.kind == "densek":
                self._flush(run)
                run = []
                self.apply_dense(op.qubits, op.matrix)
            else:
                run.append(op)
        self._flush(run)

    def plan_flush(self, run):
        """What _flush does for a run of 1-/2-qubit ops, as a list of steps (host only): ("pass", planner.Pass) = one
        launch of the fused pass kernel, ("swap", pairs) = one qsv_swap_qubit_pairs launch.  SWAP gates are taken out of the
        stream by relabelling (planner.eliminate_swaps) and the qubit permutation they leave is settled at the end."""
        ops, rounds = eliminate_swaps(merge_ops(run), self.n)
        steps = [("pass", p) for p in plan_passes(ops, self.n, self.tile_qubits, tails=True)]
        steps += [("swap", pairs) for pairs in rounds]
        return steps

    def run_step(self, step):
        kind, what = step
        if kind == "pass":
            self.apply_pass(what.tile_qubits, what.ops, tail=what.tail)
        else:
            self.swap_qubit_pairs(what)

    def swap_qubit_pairs(self, pairs):
        """SWAP gates on disjoint qubit pairs as one in-place pass over HBM (qsv_swap_qubit_pairs)."""
        flat = [int(q) for pair in pairs for q in pair]
        with self._ctx():
            _lib.check(self.lib.qsv_swap_qubit_pairs(self.ptr, self.n, _int32_array(flat), len(pairs), self.stream),
                       "qsv_swap_qubit_pairs")
        self.passes_launched += 1
        self.gates_applied += len(pairs)

    def _flush(self, run):
        if not run:
            return
        for step in self.plan_flush(run):
            self.run_step(step)

    # -- measurement -----------------------------------------------------------------
    def prob_qubit(self, qubit):
        out = (ctypes.c_double * 2)()
        with self._ctx():
            _lib.check(self.lib.qsv_prob_qubit(self.ptr, self.n, int(qubit),
                                               ctypes.c_void_p(self._reduce_ws.data_ptr()), out, self.stream),
                       "qsv_prob_qubit")
        return float(out[0]), float(out[1])

    def norm2(self):
        out = ctypes.c_double(0.0)
        with self._ctx():
            _lib.check(self.lib.qsv_norm2(self.ptr, self.n, ctypes.c_void_p(self._reduce_ws.data_ptr()),
                                          ctypes.byref(out), self.stream), "qsv_norm2")
        return float(out.value)

    def expval_pauli(self, pauli):
        """<psi|P|psi> for a Pauli label such as "XIZY" (leftmost character = highest qubit, as in
        qiskit.quantum_info.Pauli labels).  Mirrors Statevector._expectation_value_pauli
        (quantum_info/states/statevector.py:377-408): masks + phase on the host, the 2^n loop on the GPU."""
        label = pauli.upper()
        if len(label) != self.n or any(c not in "IXYZ" for c in label):
            raise ValueError("Pauli label must have %d characters from IXYZ" % self.n)
        z_mask = x_mask = 0
        n_y = 0
        for q, c in enumerate(reversed(label)):
            if c in "ZY":
                z_mask |= 1 << q
            if c in "XY":
                x_mask |= 1 << q
            n_y += c == "Y"
        if z_mask == 0 and x_mask == 0:
            return float(np.sqrt(self.norm2()))  # the reference returns the norm here (statevector.py:397-398)
        phase = (-1j) ** n_y  # statevector.py:403-404
        out = ctypes.c_double(0.0)
        with self._ctx():
            _lib.check(self.lib.qsv_expval_pauli(self.ptr, self.n, z_mask, x_mask, float(phase.real),
                                                 float(phase.imag), ctypes.c_void_p(self._reduce_ws.data_ptr()),
                                                 ctypes.byref(out), self.stream), "qsv_expval_pauli")
        return float(out.value)

    def collapse(self, qubit, outcome, prob, is_reset=False):
        with self._ctx():
            _lib.check(self.lib.qsv_collapse(self.ptr, self.n, int(qubit), int(outcome), float(prob),
                                             1 if is_reset else 0, self.stream), "qsv_collapse")
        self.passes_launched += 1
        self.gates_applied += 1

    # -- shot-batched mid-circuit operations: the top ``n_shot_qubits`` of this state number independent shots ----
    def _batch_ws(self, n_shot_qubits):
        need = int(self.lib.qsv_batch_workspace_bytes(n_shot_qubits)) // 8 + 8
        if getattr(self, "_batch_scratch", None) is None or self._batch_scratch.numel() < need:
            self._batch_scratch = self._torch.empty(need, dtype=self._torch.float64, device=self.device)
        return ctypes.c_void_p(self._batch_scratch.data_ptr())

    def collapse_batch(self, n_shot_qubits, qubit, outcomes, scales, is_reset=False):
        """Per-shot projective collapse of ``qubit``: shot s keeps outcome ``outcomes[s]`` scaled by ``scales[s]``
        (= 1/sqrt(p_s)); with ``is_reset`` the |1> half moves to |0> (qsv_collapse_batch)."""
        out = np.ascontiguousarray(outcomes, dtype=np.uint8)
        sc = np.ascontiguousarray(scales, dtype=np.float64)
        assert out.size == sc.size == 1 << n_shot_qubits
        with self._ctx():
            _lib.check(self.lib.qsv_collapse_batch(self.ptr, self.n, int(n_shot_qubits), int(qubit),
                                                   ctypes.c_void_p(out.ctypes.data), ctypes.c_void_p(sc.ctypes.data),
                                                   1 if is_reset else 0, self._batch_ws(n_shot_qubits), self.stream),
                       "qsv_collapse_batch")
        self.passes_launched += 1

    def apply_masked(self, n_shot_qubits, op, mask):
        """The 1-/2-qubit GateOp ``op`` on the shots whose ``mask`` entry is true (qsv_apply_masked)."""
        gate = _lib.QsvGate()
        _fill_gate(gate, op)
        m = np.ascontiguousarray(mask, dtype=np.uint8)
        assert m.size == 1 << n_shot_qubits
        with self._ctx():
            _lib.check(self.lib.qsv_apply_masked(self.ptr, self.n, int(n_shot_qubits), ctypes.byref(gate),
                                                 ctypes.c_void_p(m.ctypes.data), self._batch_ws(n_shot_qubits),
                                                 self.stream), "qsv_apply_masked")
        self.passes_launched += 1
        self.gates_applied += 1

    def marginal(self, measured_sorted):
        """Device tensor of 2^m marginal probabilities (bit pos <-> measured_sorted[pos])."""
        torch = self._torch
        m = len(measured_sorted)
        ws_bytes = self.lib.qsv_marginal_workspace_bytes(self.n, m)
        if self._marg_ws is None or self._marg_ws.numel() * 8 < ws_bytes:
            self._marg_ws = torch.empty(ws_bytes // 8 + 8, dtype=torch.float64, device=self.device)
        probs = torch.empty(1 << m, dtype=torch.float64, device=self.device)
        with self._ctx():
            _lib.check(self.lib.qsv_marginal(self.ptr, self.n, _int32_array(measured_sorted), m,
                                             ctypes.c_void_p(probs.data_ptr()),
                                             ctypes.c_void_p(self._marg_ws.data_ptr()), self.stream),
                       "qsv_marginal")
        return probs

    def sample(self, measured_sorted, uniforms, exact_order=True, check_norm=True):
        """searchsorted(cumsum(p)/sum(p), uniforms, 'right') over the marginal of the measured qubits.
        Returns (int64 ndarray of bin indices, total probability)."""
        m = len(measured_sorted)
        if m == self.n:
            return self._sample_from(self.state, 1, self.n, uniforms, exact_order, check_norm)
        self._probs = self.marginal(measured_sorted)
        return self._sample_from(self._probs, 0, m, uniforms, exact_order, check_norm)

    def sample_probs(self, probs, log2_bins, uniforms, exact_order=True, check_norm=True):
        """The same sampler over a device tensor of 2^log2_bins probabilities (the all-reduced marginal of a
        sharded state, dist.ShardedState.sample)."""
        return self._sample_from(probs, 0, int(log2_bins), uniforms, exact_order, check_norm)

    def _sample_from(self, src, from_amps, log2_bins, uniforms, exact_order, check_norm):
        torch = self._torch
        uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
        shots = int(uniforms.size)
        ws_bytes = self.lib.qsv_sample_workspace_bytes(log2_bins, shots)
        if self._sample_ws is None or self._sample_ws.numel() * 8 < ws_bytes:
            self._sample_ws = torch.empty(ws_bytes // 8 + 8, dtype=torch.float64, device=self.device)
        out = np.empty(shots, dtype=np.int64)
        total = ctypes.c_double(0.0)
        with self._ctx():
            _lib.check(self.lib.qsv_sample(ctypes.c_void_p(src.data_ptr()), from_amps, log2_bins,
                                           ctypes.c_void_p(uniforms.ctypes.data), shots,
                                           ctypes.c_void_p(out.ctypes.data),
                                           (1 if exact_order else 0) | (0 if check_norm else 2),
                                           ctypes.c_void_p(self._sample_ws.data_ptr()), ctypes.byref(total),
                                           self.stream), "qsv_sample")
        return out, float(total.value)

    # -- the exact-order sampler phase by phase, over ALL qubits of this state (dist.ShardedState.sample_all runs
    # -- them across the shards: the running sum of shard r starts where shard r - 1 ended) ------------------
    def _cdf_ws(self, shots):
        ws_bytes = self.lib.qsv_sample_workspace_bytes(self.n, int(shots))
        if self._sample_ws is None or self._sample_ws.numel() * 8 < ws_bytes:
            self._sample_ws = self._torch.empty(ws_bytes // 8 + 8, dtype=self._torch.float64, device=self.device)
        return ctypes.c_void_p(self._sample_ws.data_ptr())

    def cdf_prepare(self, shots):
        """Blocked chunk sums; returns the (approximate) sum of |psi|^2 over this state."""
        out = ctypes.c_double(0.0)
        with self._ctx():
            _lib.check(self.lib.qsv_cdf_prepare(self.ptr, 1, self.n, self._cdf_ws(shots), ctypes.byref(out),
                                                self.stream), "qsv_cdf_prepare")
        return float(out.value)

    def cdf_classify(self, approx_offset, shots):
        with self._ctx():
            _lib.check(self.lib.qsv_cdf_classify(self.ptr, 1, self.n, float(approx_offset), self._cdf_ws(shots),
                                                 self.stream), "qsv_cdf_classify")

    def cdf_chain(self, exact_start, shots):
        """numpy-cumsum running sum over this state starting from ``exact_start``; returns its last value."""
        out = ctypes.c_double(0.0)
        with self._ctx():
            _lib.check(self.lib.qsv_cdf_chain(self.ptr, 1, self.n, float(exact_start), self._cdf_ws(shots),
                                              ctypes.byref(out), self.stream), "qsv_cdf_chain")
        return float(out.value)

    def cdf_search(self, exact_start, total, uniforms):
        """Bin index per uniform, -1 / -2 for shots that fall before / after this state's share of the CDF."""
        uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
        out = np.empty(uniforms.size, dtype=np.int64)
        with self._ctx():
            _lib.check(self.lib.qsv_cdf_search(self.ptr, 1, self.n, float(exact_start), float(total),
                                               ctypes.c_void_p(uniforms.ctypes.data), int(uniforms.size),
                                               ctypes.c_void_p(out.ctypes.data), self._cdf_ws(uniforms.size),
                                               self.stream), "qsv_cdf_search")
        return out

    # -- readout ---------------------------------------------------------------------
    def chop(self, threshold):
        with self._ctx():
            _lib.check(self.lib.qsv_chop(self.ptr, self.n, float(threshold), self.stream), "qsv_chop")

    def download(self):
        """Host copy (a fresh complex128 ndarray owned by the caller) of the amplitudes."""
        host = np.empty(1 << self.n, dtype=np.complex128)
        with self._ctx():
            _lib.check(self.lib.qsv_download(self.ptr, self.n, ctypes.c_void_p(host.ctypes.data), self.stream),
                       "qsv_download")
        return host

    # -- multi-GPU staging (dist.py) ---------------------------------------------------
    def comm_device(self):
        return self.device

    def pack_half(self, local_qubit, bit_value, buf, begin, count):
        """buf[0:2*count] <- the amplitudes [begin, begin+count) of the half with bit local_qubit == bit_value."""
        with self._ctx():
            _lib.check(self.lib.qsv_pack_half(self.ptr, self.n, int(local_qubit), int(bit_value),
                                              ctypes.c_void_p(buf.data_ptr()), int(begin), int(count),
                                              self.stream), "qsv_pack_half")

    def unpack_half(self, local_qubit, bit_value, buf, begin, count):
        with self._ctx():
            _lib.check(self.lib.qsv_unpack_half(self.ptr, self.n, int(local_qubit), int(bit_value),
                                                ctypes.c_void_p(buf.data_ptr()), int(begin), int(count),
                                                self.stream), "qsv_unpack_half")

    def pack_bits(self, local_qubits, pattern, buf, begin, count):
        """buf <- amplitudes [begin, begin+count) of the sub-block with local_qubits[i] == bit i of pattern."""
        with self._ctx():
            _lib.check(self.lib.qsv_pack_bits(self.ptr, self.n, _int32_array(local_qubits), len(local_qubits),
                                              int(pattern), ctypes.c_void_p(buf.data_ptr()), int(begin), int(count),
                                              self.stream), "qsv_pack_bits")

    def unpack_bits(self, local_qubits, pattern, buf, begin, count):
        with self._ctx():
            _lib.check(self.lib.qsv_unpack_bits(self.ptr, self.n, _int32_array(local_qubits), len(local_qubits),
                                                int(pattern), ctypes.c_void_p(buf.data_ptr()), int(begin), int(count),
                                                self.stream), "qsv_unpack_bits")

    # -- NVLink peer memory (dist.ShardedState): CUDA IPC handle of this state, peers' states, the in-place swap ----
    def ipc_export(self):
        """(64-byte handle, offset) that lets another process on this node map this state's amplitudes."""
        handle = ctypes.create_string_buffer(64)
        offset = ctypes.c_uint64(0)
        with self._ctx():
            _lib.check(self.lib.qsv_ipc_export(self.ptr, handle, ctypes.byref(offset)), "qsv_ipc_export")
        return bytes(handle.raw), int(offset.value)

    def ipc_import(self, handle, offset):
        """Map a peer's state (from its ipc_export) into this process; returns the peer pointer (int)."""
        out = ctypes.c_void_p()
        with self._ctx():
            _lib.check(self.lib.qsv_ipc_import(ctypes.c_char_p(handle), int(offset), ctypes.byref(out)), "qsv_ipc_import")
        return int(out.value)

    def ipc_release(self, peer_ptr, offset):
        with self._ctx():
            _lib.check(self.lib.qsv_ipc_release(ctypes.c_void_p(peer_ptr), int(offset)), "qsv_ipc_release")

    def swap_bits_p2p(self, local_qubits, my_pattern, peer_ptrs, max_ctas=0):
        """In-place exchange of this state's sub-blocks with the peers' over NVLink (qsv_swap_bits_p2p);
        ``peer_ptrs[lam]`` = mapped state of the rank whose swapped global bits read ``lam`` (None for my own)."""
        arr = (ctypes.c_void_p * len(peer_ptrs))(*[ctypes.c_void_p(p) if p else None for p in peer_ptrs])
        with self._ctx():
            _lib.check(self.lib.qsv_swap_bits_p2p(self.ptr, self.n, _int32_array(local_qubits), len(local_qubits),
                                                  int(my_pattern), arr, int(max_ctas), self.stream),
                       "qsv_swap_bits_p2p")

    def scatter_bits(self, src, dest_bits, base, log2_dst):
        """Device tensor of 2^log2_dst doubles, zero except dst[base | deposit(s, dest_bits)] = src[s]
        (qsv_scatter_bits; dist.ShardedState.marginal)."""
        dst = self._torch.empty(1 << log2_dst, dtype=self._torch.float64, device=self.device)
        with self._ctx():
            _lib.check(self.lib.qsv_scatter_bits(ctypes.c_void_p(src.data_ptr()), len(dest_bits),
                                                 _int32_array(dest_bits), int(base), ctypes.c_void_p(dst.data_ptr()),
                                                 int(log2_dst), self.stream), "qsv_scatter_bits")
        return dst

    def snapshot(self):
        """Device copy of the amplitudes (one qsv_copy_d2d; the buffer comes from the caching allocator)."""
        snap = self._torch.empty_like(self.state)
        with self._ctx():
            _lib.check(self.lib.qsv_copy_d2d(ctypes.c_void_p(snap.data_ptr()), self.ptr, 16 << self.n, self.stream),
                       "qsv_copy_d2d")
        return snap

    def restore(self, snap):
        with self._ctx():
            _lib.check(self.lib.qsv_copy_d2d(self.ptr, ctypes.c_void_p(snap.data_ptr()), 16 << self.n, self.stream),
                       "qsv_copy_d2d")
