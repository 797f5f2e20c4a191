"""GPU parity tests of the libqsv kernels against the numpy oracle, through the C-ABI (ctypes)."""
import ctypes
import itertools

import numpy as np
import pytest

from oracle import basicaer_oracle as orc
from qiskit_sdk_py_b200 import _lib, lower
from qiskit_sdk_py_b200.lower import GateOp

pytestmark = pytest.mark.gpu
TOL = 1e-12


def rand_state(n, seed):
    rng = np.random.RandomState(seed)
    v = rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)
    return v / np.linalg.norm(v)


def rand_unitary(k, rng):
    z = rng.standard_normal((2 ** k, 2 ** k)) + 1j * rng.standard_normal((2 ** k, 2 ** k))
    q, r = np.linalg.qr(z)
    return q * (np.diag(r) / np.abs(np.diag(r)))


def oracle_apply(vec, op, n):
    if op.kind in (_lib.GATE_DIAG1, _lib.GATE_DIAG2):
        mat = np.diag(op.matrix)
    elif op.kind == _lib.GATE_CX:
        mat = orc.cx_gate_matrix()
    else:
        mat = op.matrix
    return orc.apply_unitary(vec.reshape([2] * n), mat, list(op.qubits)).reshape(-1)


def make_state(n, vec=None, **kw):
    from qiskit_sdk_py_b200.engine import DeviceState

    st = DeviceState(n, **kw)
    if vec is not None:
        st.upload(vec)
    return st


def random_ops(n, count, rng):
    ops = []
    for _ in range(count):
        kind = rng.randint(5)
        if n == 1 and kind in (1, 3, 4):
            kind = 0
        if kind == 0:
            ops.append(GateOp(_lib.GATE_DENSE1, [rng.randint(n)], rand_unitary(1, rng)))
        elif kind == 1:
            a, b = rng.choice(n, 2, replace=False)
            ops.append(GateOp(_lib.GATE_DENSE2, [a, b], rand_unitary(2, rng)))
        elif kind == 2:
            ops.append(GateOp(_lib.GATE_DIAG1, [rng.randint(n)], np.exp(1j * rng.uniform(0, 6, 2))))
        elif kind == 3:
            a, b = rng.choice(n, 2, replace=False)
            ops.append(GateOp(_lib.GATE_DIAG2, [a, b], np.exp(1j * rng.uniform(0, 6, 4))))
        else:
            a, b = rng.choice(n, 2, replace=False)
            ops.append(GateOp(_lib.GATE_CX, [a, b], None))
    return ops


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 9, 12, 13, 15])
def test_single_gates_every_qubit(n):
    """Each gate kind on every qubit / many pairs, one gate per pass (K1-K3)."""
    rng = np.random.RandomState(n)
    vec = rand_state(n, 100 + n)
    st = make_state(n, vec)
    ref = vec.copy()
    pairs = list(itertools.permutations(range(n), 2))
    if len(pairs) > 40:
        pairs = [pairs[i] for i in rng.choice(len(pairs), 40, replace=False)]
    ops = []
    for q in range(n):
        ops.append(GateOp(_lib.GATE_DENSE1, [q], rand_unitary(1, rng)))
        ops.append(GateOp(_lib.GATE_DIAG1, [q], np.exp(1j * rng.uniform(0, 6, 2))))
    for a, b in pairs:
        ops.append(GateOp(_lib.GATE_DENSE2, [a, b], rand_unitary(2, rng)))
        ops.append(GateOp(_lib.GATE_CX, [a, b], None))
        ops.append(GateOp(_lib.GATE_DIAG2, [a, b], np.exp(1j * rng.uniform(0, 6, 4))))
    for op in ops:
        tile = _default_tile(n, op.qubits, min(n, 12))
        st.apply_pass(tile, [op])
        ref = oracle_apply(ref, op, n)
    got = st.download()
    assert np.max(np.abs(got - ref)) < TOL


def _default_tile(n, qubits, size):
    tile = list(qubits)
    for q in range(n):
        if len(tile) >= size:
            break
        if q not in tile:
            tile.append(q)
    return sorted(tile)


@pytest.mark.parametrize("n,tile_size", [(3, 3), (6, 4), (10, 5), (10, 10), (14, 8), (14, 12), (14, 11),
                                         (16, 12), (16, 9), (18, 11)])
def test_fused_passes_any_tile(n, tile_size):
    """Runs of gates fused by the planner into cache-blocked passes with arbitrary tile qubit sets (K4)."""
    rng = np.random.RandomState(1000 + n + tile_size)
    vec = rand_state(n, 7 + n)
    st = make_state(n, vec, tile_qubits=tile_size)
    ops = random_ops(n, 60, rng)
    st.apply_ops(ops)
    ref = vec.copy()
    for op in ops:
        ref = oracle_apply(ref, op, n)
    assert st.passes_launched < len(ops)
    assert np.max(np.abs(st.download() - ref)) < TOL


@pytest.mark.parametrize("n", [8, 14])
def test_explicit_high_tiles(n):
    """Tiles made only of high qubits (no low qubits): still correct, just not coalesced."""
    rng = np.random.RandomState(5)
    vec = rand_state(n, 3)
    st = make_state(n, vec)
    ref = vec.copy()
    tile = list(range(n - 5, n))
    ops = [GateOp(_lib.GATE_DENSE2, [n - 1, n - 4], rand_unitary(2, rng)),
           GateOp(_lib.GATE_DENSE1, [n - 5], rand_unitary(1, rng)),
           GateOp(_lib.GATE_CX, [n - 2, n - 3], None)]
    st.apply_pass(tile, ops)
    for op in ops:
        ref = oracle_apply(ref, op, n)
    assert np.max(np.abs(st.download() - ref)) < TOL


@pytest.mark.parametrize("n,tile", [(9, 6), (13, 9), (15, 11), (16, 12)])
def test_remap_passes(n, tile):
    """qsv_apply_pass_remap: gates + write-back qubit permutation, then plan_restore brings the layout back."""
    from qiskit_sdk_py_b200 import circuits
    from qiskit_sdk_py_b200.planner import plan_passes, plan_restore

    ins = circuits.random_basis_instructions(n, 10, 9) + circuits.quantum_volume_instructions(n, 4, 5)
    ops = [o for o in (lower.lower_gate(i["name"], i["qubits"], i.get("params")) for i in ins) if o is not None]
    vec = rand_state(n, 11)
    st = make_state(n, vec)
    ref = vec.copy()
    for op in ops:
        ref = oracle_apply(ref, op, n)
    layout = list(range(n))
    passes = plan_passes(ops, n, tile, layout=layout, remap=True)
    assert any(p.dest for p in passes)
    for p in passes + plan_restore(layout, n, tile):
        st.apply_pass(p.tile_qubits, p.ops, p.dest)
    assert layout == list(range(n))
    assert np.max(np.abs(st.download() - ref)) < TOL


def _random_tail(n, tile, count, rng):
    """Diagonal gates allowed in a pass tail: 1-qubit anywhere, 2-qubit with at least one qubit outside the tile."""
    outside = [q for q in range(n) if q not in tile]
    tail = []
    for _ in range(count):
        kind = rng.randint(4)
        if kind == 0 or not outside:
            tail.append(GateOp(_lib.GATE_DIAG1, [rng.randint(n)], np.exp(1j * rng.uniform(0, 6, 2))))
            continue
        a = outside[rng.randint(len(outside))]
        if kind == 3 and len(outside) >= 2:
            b = a
            while b == a:
                b = outside[rng.randint(len(outside))]
        else:
            b = tile[rng.randint(len(tile))]
        qs = [a, b] if rng.randint(2) else [b, a]
        tail.append(GateOp(_lib.GATE_DIAG2, qs, np.exp(1j * rng.uniform(0, 6, 4))))
    return tail


@pytest.mark.parametrize("n,tile_size,n_ops,n_tail", [(5, 3, 3, 10), (9, 6, 0, 40), (12, 8, 6, 1), (14, 11, 10, 192),
                                                      (16, 12, 5, 100), (13, 4, 2, 60), (15, 11, 0, 7)])
def test_pass_with_diagonal_tail(n, tile_size, n_ops, n_tail):
    """qsv_apply_pass_tail: diagonal gates on qubits outside the tile, applied inside the pass (per tile they
    collapse to factor pairs on the tile bits; the kernel rebuilds two small tables per tile)."""
    rng = np.random.RandomState(31 * n + n_tail)
    vec = rand_state(n, 5)
    tile = sorted(int(q) for q in rng.choice(n, tile_size, replace=False))
    if tile_size >= 8:
        tile = sorted(set([0, 1, 2, 3]) | set(tile[4:]))
        tile = tile + [q for q in range(n) if q not in tile][: tile_size - len(tile)]
    ops = []
    for _ in range(n_ops):
        a, b = (int(x) for x in rng.choice(tile, 2, replace=False))
        pick = rng.randint(3)
        if pick == 0:
            ops.append(GateOp(_lib.GATE_DENSE2, [a, b], rand_unitary(2, rng)))
        elif pick == 1:
            ops.append(GateOp(_lib.GATE_DIAG2, [a, b], np.exp(1j * rng.uniform(0, 6, 4))))
        else:
            ops.append(GateOp(_lib.GATE_DENSE1, [a], rand_unitary(1, rng)))
    tail = _random_tail(n, tile, n_tail, rng)
    st = make_state(n, vec)
    st.apply_pass(tile, ops, tail=tail)
    ref = vec.copy()
    for op in ops + tail:
        ref = oracle_apply(ref, op, n)
    assert np.max(np.abs(st.download() - ref)) < TOL


@pytest.mark.parametrize("n,seed,n_tail", [(22, 1, 0), (22, 2, 25), (23, 3, 0), (22, 4, 60)])
def test_group_kernel_passes(n, seed, n_tail):
    """Passes with an 11-qubit tile and >= 2048 tiles run on the three-group double-buffered variant of the pass
    kernel (tile_pass_kernel<864,1,3,true>: 8 compute warps + a TMA producer warp per group): same result as the
    oracle, and bit for bit the result of the variant without producer warps (flag 512) and of the single-buffer
    variant (flag 16), whose arithmetic per tile is identical."""
    from qiskit_sdk_py_b200.engine import DeviceState

    rng = np.random.RandomState(seed)
    vec = rand_state(n, 40 + seed)
    plan = []
    for _ in range(3):
        tile = [0, 1, 2, 3] + sorted(int(q) for q in rng.choice(np.arange(4, n), 7, replace=False))
        ops = []
        for _ in range(int(rng.randint(1, 9))):
            a, b = (int(x) for x in rng.choice(tile, 2, replace=False))
            pick = rng.randint(5)
            if pick <= 1:
                ops.append(GateOp(_lib.GATE_DENSE2, [a, b], rand_unitary(2, rng)))
            elif pick == 2:
                ops.append(GateOp(_lib.GATE_DIAG2, [a, b], np.exp(1j * rng.uniform(0, 6, 4))))
            elif pick == 3:
                ops.append(GateOp(_lib.GATE_DENSE1, [a], rand_unitary(1, rng)))
            else:
                ops.append(GateOp(_lib.GATE_CX, [a, b], None))
        plan.append((tile, ops, _random_tail(n, tile, n_tail, rng)))
    plan.append(([0, 1, 2, 3] + list(range(n - 7, n)), [], []))  # copy-only pass
    results = []
    lib = _lib.load()
    try:
        for flags, threads in ((0, 864), (512, 768), (16, 256)):
            assert lib.qsv_debug_set_flags(flags) == 0
            for tile, ops, tail in plan:
                assert DeviceState.pass_info(n, tile, ops)["threads"] == (threads if flags != 512 else 864)
            st = make_state(n, vec)
            for tile, ops, tail in plan:
                st.apply_pass(tile, ops, tail=tail or None)
            results.append(st.download())
            del st
    finally:
        lib.qsv_debug_set_flags(0)
    ref = vec.copy()
    for tile, ops, tail in plan:
        for op in ops + tail:
            ref = oracle_apply(ref, op, n)
    assert np.max(np.abs(results[0] - ref)) < TOL
    assert np.array_equal(results[0], results[1])
    assert np.array_equal(results[0], results[2])


@pytest.mark.parametrize("cover", [9, 10, 11])
def test_group_kernel_stage_boundaries(cover):
    """Passes of two and more stages on the warp-group variant: at a stage boundary only the warps that agree on the
    outer bits both stages share exchange data, so the kernel synchronises just those 2 or 4 warps (PassStage::sync_mask;
    gates on 9 / 10 / 11 of the tile qubits leave 2 / 1 / 0 shared outer bits).  Same amplitudes as the oracle and, bit
    for bit, as the kernel with a full group barrier at every boundary (flag 2048) and as the single-buffer variant."""
    from qiskit_sdk_py_b200.engine import DeviceState

    n = 22
    rng = np.random.RandomState(100 + cover)
    vec = rand_state(n, 70 + cover)
    plan = []
    for rep in range(3):
        tile = [0, 1, 2, 3] + sorted(int(q) for q in rng.choice(np.arange(4, n), 7, replace=False))
        t = [tile[i] for i in rng.permutation(11)]
        pairs = [(t[0], t[1]), (t[2], t[3]), (t[4], t[5]), (t[6], t[7]), (t[8], t[0])]
        if cover >= 10:
            pairs.append((t[9], t[1]))
        if cover >= 11:
            pairs.append((t[10], t[2]))
        ops = [GateOp(_lib.GATE_DENSE2, [a, b], rand_unitary(2, rng)) for a, b in pairs]  # need `cover` > 8 inner bits
        for a, b in [(t[1], t[2]), (t[8], t[3]), (t[5], t[0])]:
            ops.append(GateOp(_lib.GATE_DENSE2, [a, b], rand_unitary(2, rng)) if rng.randint(3) else
                       GateOp(_lib.GATE_DIAG2, [a, b], np.exp(1j * rng.uniform(0, 6, 4))))
        assert DeviceState.pass_info(n, tile, ops)["stages"] >= 2
        plan.append((tile, ops))
    results = []
    lib = _lib.load()
    try:
        for flags in (0, 2048, 16):
            assert lib.qsv_debug_set_flags(flags) == 0
            st = make_state(n, vec)
            for tile, ops in plan:
                st.apply_pass(tile, ops)
            results.append(st.download())
            del st
    finally:
        lib.qsv_debug_set_flags(0)
    ref = vec.copy()
    for tile, ops in plan:
        for op in ops:
            ref = oracle_apply(ref, op, n)
    assert np.max(np.abs(results[0] - ref)) < TOL
    assert np.array_equal(results[0], results[1])
    assert np.array_equal(results[0], results[2])


def test_tail_argument_errors():
    st = make_state(10, rand_state(10, 1))
    tile = list(range(8))
    diag2 = lambda a, b: GateOp(_lib.GATE_DIAG2, [a, b], np.ones(4, dtype=complex))
    with pytest.raises(_lib.QsvLibraryError, match="needs a qubit outside the tile"):
        st.apply_pass(tile, [], tail=[diag2(1, 2)])
    with pytest.raises(_lib.QsvLibraryError, match="must be diagonal"):
        st.apply_pass(tile, [], tail=[GateOp(_lib.GATE_DENSE1, [9], np.eye(2, dtype=complex))])
    with pytest.raises(_lib.QsvLibraryError, match="too many tail gates"):
        st.apply_pass(tile, [], tail=[diag2(1, 9)] * 193)


@pytest.mark.parametrize("n", [12, 17, 20])
def test_qft_through_planner_tails(n):
    """QFT in the simulator basis: merge_ops + tails bring it down to a handful of passes; same state as the
    oracle applying the 5 n (n-1) / 2 + n + 3 (n // 2) basis gates one by one."""
    from qiskit_sdk_py_b200 import circuits

    ins = circuits.qft(n, state_seed=3)["experiments"][0]["instructions"]
    ops = [o for o in (lower.lower_gate(i["name"], i["qubits"], i.get("params")) for i in ins) if o is not None]
    st = make_state(n)
    st.init_basis0()
    st.apply_ops(ops)
    assert st.passes_launched <= 10
    ref = np.zeros(2 ** n, dtype=complex)
    ref[0] = 1
    for op in ops:
        ref = oracle_apply(ref, op, n)
    assert np.max(np.abs(st.download() - ref)) < 1e-11


@pytest.mark.parametrize("n,pairs", [(2, [(0, 1)]), (6, [(4, 5)]), (10, [(0, 9), (3, 4)]), (14, [(4, 13), (5, 12), (6, 11)]),
                                     (18, [(4, 17), (9, 10), (1, 16)]), (21, [(20 - j, 4 + j) for j in range(8)]),
                                     (22, [(5, 21), (4, 6)])])
def test_swap_qubit_pairs(n, pairs):
    """qsv_swap_qubit_pairs: SWAP gates on disjoint qubit pairs as one in-place pass -- bit for bit the state with those
    index bits exchanged (the reference applies a swap as three cx einsum passes, qasm_simulator.py:545-552: exact
    permutations either way)."""
    vec = rand_state(n, 300 + n)
    st = make_state(n, vec)
    st.swap_qubit_pairs(pairs)
    t = vec.reshape([2] * n)
    for a, b in pairs:
        t = np.swapaxes(t, n - 1 - a, n - 1 - b)
    assert np.array_equal(st.download(), np.ascontiguousarray(t).reshape(-1))
    st.swap_qubit_pairs(pairs)  # an involution
    assert np.array_equal(st.download(), vec)
    lib = _lib.load()
    bad = (ctypes.c_int32 * 4)(0, 1, 1, 2) if n > 2 else (ctypes.c_int32 * 4)(0, 0, 1, 1)
    assert lib.qsv_swap_qubit_pairs(st.ptr, n, bad, 2, None) != 0  # overlapping pairs are refused


def test_swaps_leave_the_gate_stream():
    """DeviceState.apply_ops takes SWAP gates (three cx) out of the stream by relabelling the later gates and settles the
    qubit permutation with qsv_swap_qubit_pairs (planner.eliminate_swaps): same state as the oracle applying every cx,
    with fewer launches than gates-only planning."""
    from qiskit_sdk_py_b200 import planner

    n = 16
    rng = np.random.RandomState(77)
    vec = rand_state(n, 78)
    ops = []
    for layer in range(4):
        for _ in range(5):
            a, b = (int(x) for x in rng.choice(n, 2, replace=False))
            ops.append(GateOp(_lib.GATE_DENSE2, [a, b], rand_unitary(2, rng)))
        for _ in range(4):
            a, b = (int(x) for x in rng.choice(np.arange(4, n), 2, replace=False))
            ops += [GateOp(_lib.GATE_CX, [a, b], None), GateOp(_lib.GATE_CX, [b, a], None), GateOp(_lib.GATE_CX, [a, b], None)]
        ops.append(GateOp(_lib.GATE_DIAG2, [int(rng.randint(n // 2)), int(n // 2 + rng.randint(n // 2))],
                          np.exp(1j * rng.uniform(0, 6, 4))))
    st = make_state(n, vec)
    steps = st.plan_flush(ops)
    assert any(kind == "swap" for kind, _ in steps)
    assert not any(planner.is_swap(o) and min(o.qubits) >= 4 for kind, p in steps if kind == "pass" for o in p.ops)
    st.apply_ops(ops)
    ref = vec.copy()
    for op in ops:
        ref = oracle_apply(ref, op, n)
    assert np.max(np.abs(st.download() - ref)) < TOL


@pytest.mark.parametrize("n,k", [(3, 3), (5, 3), (9, 4), (12, 5), (13, 6), (14, 7), (9, 9), (15, 3), (11, 4), (16, 4),
                                 (20, 5), (18, 6), (17, 7), (22, 3), (12, 8)])
def test_dense_k(n, k):
    """k-qubit `unitary` instructions, k >= 3 (qasm_simulator.py:543-546): k = 3 is one fused block of the pass
    kernel, 4 <= k <= 7 at n >= 11 the DMMA complex-GEMM kernel (dense_k.cu), the rest the scalar staged kernel."""
    rng = np.random.RandomState(n * 31 + k)
    vec = rand_state(n, n + k)
    st = make_state(n, vec)
    ref = vec.copy()
    for _ in range(3):
        qubits = [int(q) for q in rng.choice(n, k, replace=False)]
        mat = rand_unitary(k, rng)
        st.apply_ops([lower.matrix_op(mat, qubits)])
        ref = orc.apply_unitary(ref.reshape([2] * n), mat, qubits).reshape(-1)
    assert np.max(np.abs(st.download() - ref)) < TOL


def test_init_scale_chop_identity():
    st = make_state(10)
    st.init_basis0(np.exp(0.3j))
    got = st.download()
    assert got[0] == np.exp(0.3j) and not np.any(got[1:])
    st.scale(2j)
    assert abs(st.download()[0] - 2j * np.exp(0.3j)) < 1e-15
    vec = rand_state(10, 1)
    vec[5] = 1e-16
    vec[6] = 1e-16j
    vec[7] = 0.9e-15 + 0.1e-15j
    st.upload(vec)
    st.chop(1e-15)
    exp = vec.copy()
    exp[abs(exp) < 1e-15] = 0.0
    assert np.array_equal(st.download(), exp)
    st = make_state(8)
    st.init_identity(4, 1.0 + 0.0j)
    assert np.array_equal(st.download().reshape(16, 16), np.eye(16))


@pytest.mark.parametrize("n", [1, 2, 5, 10, 16, 21])
def test_prob_collapse(n):
    """_get_measure_outcome / _add_qasm_measure / _add_qasm_reset arithmetic (K7, K8)."""
    vec = rand_state(n, 50 + n)
    st = make_state(n, vec)
    p = np.abs(vec) ** 2
    idx = np.arange(2 ** n)
    for q in sorted({0, n // 2, n - 1}):
        p0, p1 = st.prob_qubit(q)
        mask = ((idx >> q) & 1).astype(bool)
        assert abs(p0 - p[~mask].sum()) < 1e-13 and abs(p1 - p[mask].sum()) < 1e-13
    assert abs(st.norm2() - 1.0) < 1e-13
    q = n // 2
    for outcome, is_reset in ((0, False), (1, False), (0, True), (1, True)):
        st.upload(vec)
        probs = st.prob_qubit(q)
        st.collapse(q, outcome, probs[outcome], is_reset)
        s = 1 / np.sqrt(probs[outcome])
        if is_reset and outcome == 1:
            mat = [[0, s], [0, 0]]
        elif outcome == 0:
            mat = [[s, 0], [0, 0]]
        else:
            mat = [[0, 0], [0, s]]
        ref = orc.apply_unitary(vec.reshape([2] * n), mat, [q]).reshape(-1)
        assert np.max(np.abs(st.download() - ref)) < TOL
        assert abs(st.norm2() - 1.0) < 1e-12


@pytest.mark.parametrize("n,measured", [(3, [0, 1, 2]), (6, [1, 4]), (10, [0]), (10, [9]), (12, [0, 3, 5, 11]),
                                        (14, list(range(14))), (14, list(range(1, 14))), (18, [2, 17]), (16, [q for q in range(16) if q not in (3, 9)]),
                                        (20, list(range(20)))])
def test_marginal_and_exact_sampling(n, measured):
    """_add_sample_measure (qasm_simulator.py:184-213): same indices as numpy's choice()."""
    vec = rand_state(n, 77 + n)
    st = make_state(n, vec)
    m = len(measured)
    axis = list(range(n))
    for q in reversed(measured):
        axis.remove(n - 1 - q)
    p = np.reshape(np.sum(np.abs(vec.reshape([2] * n)) ** 2, axis=tuple(axis)), 2 ** m)
    if m < n:
        probs = st.marginal(measured).cpu().numpy()
        assert np.max(np.abs(probs - p)) < 1e-14
    shots = 4096
    rs = np.random.RandomState(123)
    expected = rs.choice(range(2 ** m), shots, p=p)
    u = np.random.RandomState(123).random_sample(shots)
    got, total = st.sample(measured, u, exact_order=True)
    assert abs(total - 1.0) < 1e-12
    mism = int(np.sum(got != expected))
    # identical unless a uniform falls within ~1e-16 of a CDF step (amplitude rounding only)
    assert mism == 0
    fast, total2 = st.sample(measured, u, exact_order=False)
    assert abs(total2 - 1.0) < 1e-12
    assert np.mean(fast != expected) < 2e-3


def test_sampling_edge_cases():
    # basis state: every sample is that index; zero-probability plateaus are skipped like searchsorted
    n = 12
    vec = np.zeros(2 ** n, dtype=complex)
    vec[1234] = 1.0
    st = make_state(n, vec)
    u = np.random.RandomState(0).random_sample(100)
    got, _ = st.sample(list(range(n)), u)
    assert np.all(got == 1234)
    # not normalised -> numpy's ValueError
    st.scale(1.1)
    with pytest.raises(ValueError, match="probabilities do not sum to 1"):
        st.sample(list(range(n)), u)


def test_bad_arguments_fail_loudly():
    st = make_state(6, rand_state(6, 0))
    with pytest.raises(_lib.QsvLibraryError):
        st.apply_pass([0, 1, 2], [GateOp(_lib.GATE_DENSE1, [5], np.eye(2))])  # gate qubit not in tile
    with pytest.raises(_lib.QsvLibraryError):
        st.apply_pass([0, 1, 1], [])
    with pytest.raises(_lib.QsvLibraryError):
        st.apply_pass(list(range(6)) + [7], [])


def test_expval_pauli_matches_reference_golden():
    """qsv_expval_pauli vs the values the reference's Cython exp_value.pyx produced."""
    import golden_io

    fix = golden_io.load_case("expval_pauli")
    for case in fix["cases"]:
        st = make_state(case["n"], case["state"])
        for label, want in zip(case["labels"], case["values"]):
            got = st.expval_pauli(label)
            assert abs(got - want.real) < 1e-12 and abs(want.imag) < 1e-12, (case["n"], label, got, want)
    # a larger state against the oracle
    n = 20
    vec = rand_state(n, 5)
    st = make_state(n, vec)
    rng = np.random.RandomState(3)
    for _ in range(6):
        label = "".join(rng.choice(list("IXYZ"), n))
        assert abs(st.expval_pauli(label) - orc.expval_pauli(vec, label)) < 1e-12


def _sample_probs(probs, uniforms, exact=True):
    """Call qsv_sample on an explicit probability vector; returns (indices, total, chunk_end array)."""
    import ctypes

    import torch

    lib = _lib.load()
    log2_bins = int(np.log2(probs.size))
    dev = torch.as_tensor(probs, dtype=torch.float64, device="cuda")
    ws_bytes = lib.qsv_sample_workspace_bytes(log2_bins, uniforms.size)
    ws = torch.zeros(ws_bytes // 8 + 8, dtype=torch.float64, device="cuda")
    out = np.empty(uniforms.size, dtype=np.int64)
    total = ctypes.c_double(0.0)
    u = np.ascontiguousarray(uniforms, dtype=np.float64)
    rc = lib.qsv_sample(ctypes.c_void_p(dev.data_ptr()), 0, log2_bins, ctypes.c_void_p(u.ctypes.data), u.size,
                        ctypes.c_void_p(out.ctypes.data), (1 if exact else 0) | 2, ctypes.c_void_p(ws.data_ptr()),
                        ctypes.byref(total), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "qsv_sample")
    nchunks = max(1, probs.size // 1024)
    return out, total.value, ws[:nchunks].cpu().numpy()


@pytest.mark.parametrize("kind", ["random", "wide_range", "sparse", "uniform", "ties", "tiny_tail", "leading_zeros"])
def test_exact_cdf_is_bitwise_numpy_cumsum(kind):
    """The parallel exact-order scan (integer sums inside a binade) must reproduce numpy's sequential
    float64 cumsum bit for bit at every chunk boundary, and searchsorted's indices exactly."""
    rng = np.random.RandomState(hash(kind) % 1000)
    nbins = 1 << 18
    if kind == "random":
        p = rng.random_sample(nbins)
    elif kind == "wide_range":
        p = np.exp(rng.uniform(-60, 0, nbins))
    elif kind == "sparse":
        p = np.where(rng.random_sample(nbins) < 0.01, rng.random_sample(nbins), 0.0)
    elif kind == "uniform":
        p = np.full(nbins, 1.0)
    elif kind == "ties":
        p = np.ldexp(1.0, rng.randint(-70, 0, nbins))  # powers of two: exact half-way cases occur
    elif kind == "tiny_tail":
        p = np.concatenate([rng.random_sample(nbins // 2), 1e-25 * rng.random_sample(nbins // 2)])
    else:
        p = np.concatenate([np.zeros(nbins // 2 + 777), rng.random_sample(nbins // 2 - 777)])
    p = p / p.sum()
    cdf = p.cumsum()
    u = np.random.RandomState(5).random_sample(3000)
    got, total, chunk_end = _sample_probs(p, u, exact=True)
    assert total == cdf[-1]
    assert np.array_equal(chunk_end, cdf[1023::1024]), "chunk-end running sums differ from numpy's cumsum"
    expected = (cdf / cdf[-1]).searchsorted(u, side="right")
    assert np.array_equal(got, expected)


def _cdf_call(name, *args):
    _lib.check(getattr(_lib.load(), name)(*args), name)


def _run_cdf_shards(src_host, from_amps, n_shards, uniforms):
    """The four phases of the exact-order sampler (qsv_cdf_*) over `n_shards` consecutive pieces of one array, as
    dist.ShardedState.sample_all drives them across ranks; returns (bin indices, total, all chunk ends)."""
    import ctypes

    import torch

    lib = _lib.load()
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    size = src_host.size // n_shards
    log2 = int(np.log2(size))
    u = np.ascontiguousarray(uniforms, dtype=np.float64)
    devs, wss, approx = [], [], []
    for r in range(n_shards):
        part = np.ascontiguousarray(src_host[r * size:(r + 1) * size])
        dev = torch.as_tensor(part.view(np.float64), device="cuda")
        ws = torch.zeros(lib.qsv_sample_workspace_bytes(log2, u.size) // 8 + 8, dtype=torch.float64, device="cuda")
        tot = ctypes.c_double(0.0)
        _cdf_call("qsv_cdf_prepare", ctypes.c_void_p(dev.data_ptr()), from_amps, log2, ctypes.c_void_p(ws.data_ptr()),
                  ctypes.byref(tot), stream)
        devs.append(dev), wss.append(ws), approx.append(tot.value)
    starts, end = [], 0.0
    for r in range(n_shards):
        _cdf_call("qsv_cdf_classify", ctypes.c_void_p(devs[r].data_ptr()), from_amps, log2, float(sum(approx[:r])),
                  ctypes.c_void_p(wss[r].data_ptr()), stream)
        starts.append(end)
        out = ctypes.c_double(0.0)
        _cdf_call("qsv_cdf_chain", ctypes.c_void_p(devs[r].data_ptr()), from_amps, log2, end,
                  ctypes.c_void_p(wss[r].data_ptr()), ctypes.byref(out), stream)
        end = out.value
    total = end
    result = np.full(u.size, -5, dtype=np.int64)
    claimed = np.zeros(u.size, dtype=np.int64)
    for r in range(n_shards):
        idx = np.empty(u.size, dtype=np.int64)
        _cdf_call("qsv_cdf_search", ctypes.c_void_p(devs[r].data_ptr()), from_amps, log2, starts[r], total,
                  ctypes.c_void_p(u.ctypes.data), u.size, ctypes.c_void_p(idx.ctypes.data),
                  ctypes.c_void_p(wss[r].data_ptr()), stream)
        mine = idx >= 0
        result[mine] = idx[mine] + r * size
        claimed += mine
    nchunks = max(1, size // 1024)
    chunk_end = np.concatenate([w[:nchunks].cpu().numpy() for w in wss])
    assert np.all(claimed == 1)
    return result, total, chunk_end


@pytest.mark.parametrize("scale", [1.0, 1e-5])
def test_exact_cdf_from_amplitudes_is_numpy_abs2(scale):
    """All qubits measured: p_i = np.abs(psi_i) ** 2 as numpy rounds it (max * sqrt(fma(r, r, 1)) squared, NOT
    re^2 + im^2 -- about a third of random inputs differ in the last bit) and the running sum is numpy's cumsum,
    so the chunk ends and the sampled bins equal the reference's _add_sample_measure bit for bit."""
    rng = np.random.RandomState(11)
    nbins = 1 << 18
    z = (rng.standard_normal(nbins) + 1j * rng.standard_normal(nbins)) * scale
    z[rng.random_sample(nbins) < 0.01] = 0.0
    z[5] = 3e-3 * scale          # purely real / purely imaginary amplitudes
    z[6] = 2e-3j * scale
    z /= np.linalg.norm(z)
    p = np.abs(z) ** 2
    naive = z.real * z.real + z.imag * z.imag
    assert np.count_nonzero(p != naive) > nbins // 10, "the two roundings should differ often (else the test is void)"
    cdf = p.cumsum()
    u = np.random.RandomState(5).random_sample(3000)
    got, total, chunk_end = _run_cdf_shards(z, 1, 1, u)
    assert total == cdf[-1]
    assert np.array_equal(chunk_end, cdf[1023::1024])
    assert np.array_equal(got, (cdf / cdf[-1]).searchsorted(u, side="right"))
    # and through the public sampler of a state
    n = 18
    st = make_state(n, z)
    samples, tot = st.sample(list(range(n)), u, exact_order=True)
    assert tot == cdf[-1] and np.array_equal(samples, got)


@pytest.mark.parametrize("kind,n_shards", [("amps", 2), ("amps", 8), ("probs_wide", 4), ("probs_sparse", 4),
                                           ("probs_empty_shard", 4)])
def test_cdf_chained_across_shards_is_one_cumsum(kind, n_shards):
    """qsv_cdf_prepare / classify / chain / search over consecutive shards, each continuing the previous shard's
    exact running sum = numpy's cumsum over the whole array (what ShardedState.sample_all does across ranks)."""
    rng = np.random.RandomState(len(kind) + n_shards)
    nbins = 1 << 17
    u = np.random.RandomState(6).random_sample(4000)
    if kind == "amps":
        z = rng.standard_normal(nbins) + 1j * rng.standard_normal(nbins)
        z /= np.linalg.norm(z)
        p, src, from_amps = np.abs(z) ** 2, z, 1
    else:
        if kind == "probs_wide":
            p = np.exp(rng.uniform(-60, 0, nbins))
        elif kind == "probs_sparse":
            p = np.where(rng.random_sample(nbins) < 0.01, rng.random_sample(nbins), 0.0)
        else:
            p = rng.random_sample(nbins)
            p[nbins // 4: nbins // 2] = 0.0  # the second of four shards holds no probability at all
        p = p / p.sum()
        src, from_amps = p, 0
    cdf = p.cumsum()
    got, total, chunk_end = _run_cdf_shards(src, from_amps, n_shards, u)
    assert total == cdf[-1]
    assert np.array_equal(chunk_end, cdf[1023::1024])
    assert np.array_equal(got, (cdf / cdf[-1]).searchsorted(u, side="right"))


def test_scatter_bits_and_device_memory_helpers():
    """qsv_scatter_bits (a shard's marginal into the full marginal), qsv_alloc / qsv_copy_d2d / qsv_free."""
    import ctypes

    import torch

    lib = _lib.load()
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rng = np.random.RandomState(2)
    src = rng.random_sample(1 << 5)
    dest_bits, base, log2_dst = [6, 0, 3, 1, 4], (1 << 2) | (1 << 7), 8
    dev_src = torch.as_tensor(src, device="cuda")
    dev_dst = torch.full((1 << log2_dst,), 7.0, dtype=torch.float64, device="cuda")
    arr = (ctypes.c_int32 * 5)(*dest_bits)
    _lib.check(lib.qsv_scatter_bits(ctypes.c_void_p(dev_src.data_ptr()), 5, arr, base, ctypes.c_void_p(dev_dst.data_ptr()),
                                    log2_dst, stream), "qsv_scatter_bits")
    want = np.zeros(1 << log2_dst)
    for s_idx in range(32):
        t_idx = base
        for i, b in enumerate(dest_bits):
            t_idx |= ((s_idx >> i) & 1) << b
        want[t_idx] = src[s_idx]
    assert np.array_equal(dev_dst.cpu().numpy(), want)
    assert lib.qsv_scatter_bits(ctypes.c_void_p(dev_src.data_ptr()), 5, (ctypes.c_int32 * 5)(6, 0, 3, 1, 2), base,
                                ctypes.c_void_p(dev_dst.data_ptr()), log2_dst, stream) == _lib.QSV_EINVAL
    ptr = ctypes.c_void_p()
    _lib.check(lib.qsv_alloc(ctypes.byref(ptr), 32 * 8), "qsv_alloc")
    _lib.check(lib.qsv_copy_d2d(ptr, ctypes.c_void_p(dev_src.data_ptr()), 32 * 8, stream), "qsv_copy_d2d")
    back = torch.zeros(32, dtype=torch.float64, device="cuda")
    _lib.check(lib.qsv_copy_d2d(ctypes.c_void_p(back.data_ptr()), ptr, 32 * 8, stream), "qsv_copy_d2d")
    torch.cuda.synchronize()
    assert np.array_equal(back.cpu().numpy(), src)
    _lib.check(lib.qsv_free(ptr), "qsv_free")
    st = make_state(12, rand_state(12, 3))
    before = st.download()
    snap = st.snapshot()
    st.apply_ops([GateOp(_lib.GATE_DENSE1, [3], rand_unitary(1, rng))])
    st.restore(snap)
    assert np.array_equal(st.download(), before)


def test_dense_k_low_and_high_qubits():
    """The DMMA dense-k kernel with gate qubits among the lowest bits (they are tile bits either way), among the
    highest, and in descending order (matrix index bit j <-> qubits[j])."""
    n = 19
    rng = np.random.RandomState(77)
    vec = rand_state(n, 5)
    st = make_state(n, vec)
    ref = vec.copy()
    for qubits in ([0, 1, 2, 3], [18, 17, 16, 15, 14], [3, 18, 0, 9, 1, 12], [5, 4, 3, 2, 1, 0, 6], [2, 10, 18]):
        mat = rand_unitary(len(qubits), rng)
        st.apply_ops([lower.matrix_op(mat, qubits)])
        ref = orc.apply_unitary(ref.reshape([2] * n), mat, qubits).reshape(-1)
    assert np.max(np.abs(st.download() - ref)) < TOL


def test_device_statevector_handle():
    """the statevector simulator leaves the final state on the device and hands out a
    Statevector-like handle; expectation values / probabilities / sampling run there (qsv_expval_pauli, qsv_marginal,
    exact-order sampler), checked against the oracle on the downloaded amplitudes."""
    import copy

    from qiskit_sdk_py_b200 import circuits, simulator
    from qiskit_sdk_py_b200.statevector import DeviceStatevector

    n = 12
    qobj = circuits.random_circuit(n, 8, 4, measure_final=False, shots=1)
    res = simulator.StatevectorSimulatorB200(statevector_on_device=True).run(copy.deepcopy(qobj))
    sv = res["results"][0]["data"]["statevector"]
    assert isinstance(sv, DeviceStatevector) and sv.num_qubits == n
    ref = orc.OracleQasmSimulator(show_final_state=True).run(copy.deepcopy(qobj))["results"][0]["data"]["statevector"]
    assert np.max(np.abs(sv.to_numpy() - ref)) < TOL
    rng = np.random.RandomState(0)
    for label in ["Z" * n, "X" + "I" * (n - 1)] + ["".join(rng.choice(list("IXYZ"), n)) for _ in range(6)]:
        assert abs(sv.expectation_value(label) - orc.expval_pauli(ref, label)) < 1e-12
    p = np.abs(ref) ** 2
    got = sv.probabilities([7, 2, 9])
    want = np.zeros(8)
    for i in range(1 << n):
        want[((i >> 7) & 1) | (((i >> 2) & 1) << 1) | (((i >> 9) & 1) << 2)] += p[i]
    assert np.max(np.abs(got - want)) < 1e-12
    own = np.abs(sv.to_numpy()) ** 2
    cdf = own.cumsum()
    exp = (cdf / cdf[-1]).searchsorted(np.random.RandomState(9).random_sample(300), side="right")
    assert sv.sample_memory(300, seed=9) == [format(int(s), "0%db" % n) for s in exp]
