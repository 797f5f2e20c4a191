"""CPU tests of the product's host logic (lowering, planner, control flow, RNG order, result layout)
with the device replaced by tests/fake_state.FakeState (numpy oracle).  The CUDA path itself is
covered by the ``-m gpu`` tests."""
import copy

import numpy as np
import pytest

import golden_io
import helpers
from fake_state import FakeState
from qiskit_sdk_py_b200 import _lib, circuits, lower, simulator
from qiskit_sdk_py_b200.planner import plan_passes

SMALL = helpers.small_cases(12)


def fake_factory(backend):
    if backend == "qasm_simulator":
        return simulator.QasmSimulatorB200(state_factory=FakeState, max_qubits=24)
    if backend == "statevector_simulator":
        return simulator.StatevectorSimulatorB200(state_factory=FakeState, max_qubits=24)
    return simulator.UnitarySimulatorB200(state_factory=FakeState, max_qubits=12)


@pytest.mark.parametrize("name", SMALL)
def test_host_logic_matches_reference_golden(name):
    case = golden_io.load_case(name)
    result = helpers.run_case_with(fake_factory, case)
    helpers.compare_result(result, case, amp_tol=1e-12)


def _key(op):
    return (op.kind, tuple(op.qubits), id(op.matrix))


def test_planner_preserves_order_on_shared_qubits():
    ops = [lower.lower_gate(i["name"], i["qubits"], i.get("params"))
           for i in circuits.random_basis_instructions(10, 30, 5)]
    ops = [o for o in ops if o is not None]
    for tile in (3, 4, 6, 10, 12):
        passes = plan_passes(ops, 10, tile_qubits=tile, pair_1q=False)  # the same op objects
        flat = [op for p in passes for op in p.ops]
        assert sorted(map(_key, flat)) == sorted(map(_key, ops))
        # per qubit, the subsequence of gates touching it is unchanged
        for q in range(10):
            assert [_key(o) for o in flat if q in o.qubits] == [_key(o) for o in ops if q in o.qubits]
        for p in passes:
            assert len(p.tile_qubits) == min(10, tile)
            assert len(set(p.tile_qubits)) == len(p.tile_qubits)
            assert len(p.ops) <= _lib.MAX_PASS_GATES
            assert sum(o.payload_doubles for o in p.ops) <= _lib.MAX_PASS_MATRIX_DOUBLES
            for o in p.ops:
                assert set(o.qubits) <= set(p.tile_qubits)
            if tile >= 6:
                assert {0, 1, 2, 3} <= set(p.tile_qubits)


def _permute_bits(vec, n, tile, dest):
    """content of physical bit tile[i] moves to physical bit dest[i]."""
    t = vec.reshape([2] * n)
    axes = list(range(n))
    for src, dst in zip(tile, dest):
        axes[n - 1 - dst] = n - 1 - src
    return np.ascontiguousarray(t.transpose(axes)).reshape(-1)


@pytest.mark.parametrize("tails", [False, True])
@pytest.mark.parametrize("n,tile", [(9, 6), (12, 8), (14, 11)])
def test_planner_remap_and_restore(n, tile, tails):
    """Passes with write-back qubit permutations (qsv_apply_pass_remap) + plan_restore, simulated
    physically with numpy: the result must equal the plain logical simulation."""
    from oracle import basicaer_oracle as orc
    from qiskit_sdk_py_b200.planner import plan_restore

    ins = circuits.random_basis_instructions(n, 12, 7) + circuits.quantum_volume_instructions(n, 4, 3)
    ops = [o for o in (lower.lower_gate(i["name"], i["qubits"], i.get("params")) for i in ins) if o is not None]
    rng = np.random.RandomState(1)
    vec0 = rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)

    def apply(vec, op):
        if op.kind in (_lib.GATE_DIAG1, _lib.GATE_DIAG2):
            mat = np.diag(op.matrix)
        elif op.kind == _lib.GATE_CX:
            mat = orc.cx_gate_matrix()
        else:
            mat = op.matrix
        return orc.apply_unitary(vec.reshape([2] * n), mat, list(op.qubits)).reshape(-1)

    ref = vec0.copy()
    for op in ops:
        ref = apply(ref, op)
    layout = list(range(n))
    passes = plan_passes(ops, n, tile, layout=layout, remap=True, tails=tails)
    assert any(p.dest for p in passes)
    if tails and tile <= 6:
        assert any(p.tail for p in passes)
    phys = vec0.copy()
    for p in passes + plan_restore(layout, n, tile):
        assert len(p.tile_qubits) == tile and {0, 1, 2, 3} <= set(p.tile_qubits)
        for op in p.ops:
            assert set(op.qubits) <= set(p.tile_qubits)
            phys = apply(phys, op)
        for op in p.tail:  # diagonal gates with a qubit outside the tile (or 1-qubit), after the pass's gates
            assert op.kind in (_lib.GATE_DIAG1, _lib.GATE_DIAG2)
            assert len(op.qubits) == 1 or not set(op.qubits) <= set(p.tile_qubits)
            phys = apply(phys, op)
        if p.dest:
            assert sorted(p.dest) == sorted(p.tile_qubits)
            phys = _permute_bits(phys, n, p.tile_qubits, p.dest)
    assert layout == list(range(n))
    assert np.max(np.abs(phys - ref)) < 1e-12


def test_planner_fuses_quantum_volume():
    n = 24
    ops = [lower.lower_gate(i["name"], i["qubits"], i.get("params"))
           for i in circuits.quantum_volume_instructions(n, n, 1234)]
    passes = plan_passes(ops, n, tile_qubits=12)
    assert len(passes) < len(ops) / 3


def test_lowering_kinds():
    assert lower.lower_gate("u1", [3], [0.5]).kind == _lib.GATE_DIAG1
    assert lower.lower_gate("rz", [3], [0.5]).kind == _lib.GATE_DIAG1
    assert lower.lower_gate("u3", [0], [0.1, 0.2, 0.3]).kind == _lib.GATE_DENSE1
    assert lower.lower_gate("cx", [1, 0], None).kind == _lib.GATE_CX
    assert lower.lower_gate("id", [1], None) is None
    assert lower.lower_gate("unitary", [0, 1], [np.diag([1, 1j, -1, -1j])]).kind == _lib.GATE_DIAG2
    assert lower.lower_gate("unitary", [0, 1, 2], [np.eye(8)]).kind == "densek"
    for name, params in (("u1", [0.3]), ("u2", [0.1, 0.2]), ("u3", [0.1, 0.2, 0.3]), ("U", [1, 2, 3]),
                         ("rz", [0.7]), ("sx", []), ("x", [])):
        from oracle import basicaer_oracle as orc
        assert np.array_equal(lower.single_gate_matrix(name, params), orc.single_gate_matrix(name, params))


def test_errors_mirror_reference_messages():
    sim = simulator.QasmSimulatorB200(state_factory=FakeState, max_qubits=24)
    qobj = circuits.bell()
    with pytest.raises(simulator.BasicAerError, match="initial statevector is not normalized"):
        sim.run(copy.deepcopy(qobj), initial_statevector=[1, 1, 0, 0])
    with pytest.raises(simulator.BasicAerError, match="initial statevector is incorrect length: 2 != 4"):
        sim.run(copy.deepcopy(qobj), initial_statevector=[1, 0])
    big = circuits.make_qobj(circuits.make_experiment("big", 50, [], 0))
    with pytest.raises(simulator.BasicAerError, match=r"Number of qubits 50 is greater than maximum \(24\)"):
        sim.run(big)
    bad = circuits.make_qobj(circuits.make_experiment("bad", 2, [{"name": "foo", "qubits": [0]}], 0))
    with pytest.raises(simulator.BasicAerError, match='qasm_simulator encountered unrecognized operation "foo"'):
        sim.run(bad)
    usim = simulator.UnitarySimulatorB200(state_factory=FakeState, max_qubits=12)
    with pytest.raises(simulator.BasicAerError, match="Unsupported"):
        usim.run(copy.deepcopy(qobj))
    with pytest.raises(simulator.BasicAerError, match="initial unitary is not unitary"):
        usim.run(circuits.make_qobj(circuits.make_experiment("u", 1, [], 0)), initial_unitary=[[1, 1], [0, 1]])


def test_random_seed_when_absent():
    sim = simulator.QasmSimulatorB200(state_factory=FakeState, max_qubits=24)
    res = sim.run(circuits.bell(seed_simulator=None))
    assert isinstance(res["results"][0]["seed_simulator"], (int, np.integer))
    assert sum(res["results"][0]["data"]["counts"].values()) == 1024


def test_get_simulator_aliases():
    assert simulator.get_simulator("local_qasm_simulator_py", state_factory=FakeState).name() == "qasm_simulator"
    with pytest.raises(LookupError):
        simulator.get_simulator("nope")


def _apply_all(ops, n, vec):
    from oracle import basicaer_oracle as orc

    for op in ops:
        if op.kind in (_lib.GATE_DIAG1, _lib.GATE_DIAG2):
            mat = np.diag(op.matrix)
        elif op.kind == _lib.GATE_CX:
            mat = orc.cx_gate_matrix()
        else:
            mat = op.matrix
        vec = orc.apply_unitary(vec.reshape([2] * n), mat, list(op.qubits)).reshape(-1)
    return vec


@pytest.mark.parametrize("seed", range(6))
def test_merge_ops_is_exact_on_random_streams(seed):
    """merge_ops changes the op list, never the state: random basis-gate streams mixed with SU(4) blocks."""
    from qiskit_sdk_py_b200.planner import merge_ops

    n = 6
    ins = circuits.random_basis_instructions(n, 40, seed) + circuits.quantum_volume_instructions(n, 3, seed)
    ins += circuits.random_basis_instructions(n, 20, seed + 50)
    ops = [o for o in (lower.lower_gate(i["name"], i["qubits"], i.get("params")) for i in ins) if o is not None]
    merged = merge_ops(ops)
    assert len(merged) < len(ops)
    rng = np.random.RandomState(seed)
    vec = rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)
    assert np.max(np.abs(_apply_all(merged, n, vec.copy()) - _apply_all(ops, n, vec.copy()))) < 1e-12
    for op in merged:
        assert len(op.qubits) <= 2 and len(set(op.qubits)) == len(op.qubits)


def test_merge_ops_turns_controlled_phase_into_one_diagonal_op():
    """QFT in the simulator basis: every cp = u1 cx u1 cx u1 collapses into a single diagonal op, so QFT-20
    (1000 instructions) becomes 200 ops: 171 diagonal 2-qubit ops and 29 dense ones (each H joins the first
    controlled phase on its qubit; the final swaps are 3 cx = one dense op)."""
    from qiskit_sdk_py_b200.planner import merge_ops

    qobj = circuits.qft(20)
    ins = qobj["experiments"][0]["instructions"]
    ops = [o for o in (lower.lower_gate(i["name"], i["qubits"], i.get("params")) for i in ins) if o is not None]
    assert len(ops) == 1000
    merged = merge_ops(ops)
    kinds = [op.kind for op in merged]
    assert len(merged) <= 230
    assert kinds.count(_lib.GATE_DIAG2) >= 170
    assert kinds.count(_lib.GATE_CX) == 0
    # untouched ops are passed through as the same objects
    lone = [lower.lower_gate("u3", [0], [0.1, 0.2, 0.3]), lower.lower_gate("cx", [1, 2], None)]
    assert merge_ops(lone)[0] is lone[0] and merge_ops(lone)[1] is lone[1]


def test_schedule_pass_ops_preserves_the_unitary():
    """planner.schedule_pass_ops only commutes gates on disjoint qubits or pairs of diagonal gates: the reordered
    list is a permutation of the input and gives the same state (oracle einsum, qasm_simulator.py:145-161)."""
    from oracle import basicaer_oracle as orc
    from qiskit_sdk_py_b200 import _lib, planner
    from qiskit_sdk_py_b200.lower import GateOp

    rng = np.random.RandomState(12)

    def ru(k):
        q, r = np.linalg.qr(rng.standard_normal((2 ** k, 2 ** k)) + 1j * rng.standard_normal((2 ** k, 2 ** k)))
        return q * (np.diag(r) / np.abs(np.diag(r)))

    n = 7
    for trial in range(30):
        ops = []
        for _ in range(int(rng.randint(3, 30))):
            a, b = (int(x) for x in rng.choice(n, 2, replace=False))
            pick = rng.randint(6)
            if pick <= 1:
                ops.append(GateOp(_lib.GATE_DENSE2, [a, b], ru(2)))
            elif pick == 2:
                ops.append(GateOp(_lib.GATE_DIAG2, [a, b], np.exp(1j * rng.uniform(0, 6, 4))))
            elif pick == 3:
                ops.append(GateOp(_lib.GATE_DIAG1, [a], np.exp(1j * rng.uniform(0, 6, 2))))
            elif pick == 4:
                ops.append(GateOp(_lib.GATE_DENSE1, [a], ru(1)))
            else:
                ops.append(GateOp(_lib.GATE_CX, [a, b], None))
        out = planner.schedule_pass_ops(ops)
        assert sorted(map(id, out)) == sorted(map(id, ops))
        vec = rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)

        def run(seq):
            st = vec.reshape([2] * n)
            for op in seq:
                mat = np.diag(op.matrix) if op.kind in (_lib.GATE_DIAG1, _lib.GATE_DIAG2) else \
                    orc.cx_gate_matrix() if op.kind == _lib.GATE_CX else op.matrix
                st = orc.apply_unitary(st, mat, list(op.qubits))
            return st.reshape(-1)

        assert np.max(np.abs(run(out) - run(ops))) < 1e-12


def test_schedule_turns_qft_fans_into_blocks_and_long_runs():
    """QFT-33 through merge_ops + plan_passes: with the in-pass scheduler the Hadamard-merged dense gates pair up into
    3-qubit blocks and the controlled-phase fans between them collapse into a few long diagonal runs."""
    from qiskit_sdk_py_b200 import circuits, lower, planner
    from qiskit_sdk_py_b200.engine import DeviceState

    n = 33
    ins = circuits.qft(n)["experiments"][0]["instructions"]
    ops = [o for o in (lower.lower_gate(i["name"], i["qubits"], i.get("params")) for i in ins) if o is not None]
    merged = planner.merge_ops(ops)
    count = {}
    for sched in (False, True):
        passes = planner.plan_passes(merged, n, tails=True, schedule=sched)
        kinds = [o["kind"] for p in passes for o in DeviceState.pass_ops(n, p.tile_qubits, p.ops)]
        count[sched] = (len(passes), kinds.count(6), kinds.count(2))
    assert count[True][0] == count[False][0] == 10  # 13 before the in-tile diagonals behind a tail joined its pass
    assert count[True][1] >= 2 * count[False][1] and count[True][2] <= count[False][2] // 2


def test_isolated_1q_gates_of_a_pass_are_paired():
    """planner.pair_isolated_1q: a layer of 1-qubit rotations costs one 2-qubit tensor op per PAIR of gates; gates whose
    qubit another gate of the pass touches are left alone (merge_ops / the kernel's blocks take care of those)."""
    from fake_state import FakeState
    from oracle import basicaer_oracle as orc
    from qiskit_sdk_py_b200 import _lib, lower, planner

    rng = np.random.RandomState(3)
    n = 9
    ops = [lower.lower_gate("u3", [q], list(rng.uniform(0, 3, 3))) for q in range(n)]
    ops.append(lower.lower_gate("cx", [0, 1], None))
    passes = planner.plan_passes(ops, n, tile_qubits=9)
    assert len(passes) == 1
    kinds = [o.kind for o in passes[0].ops]
    # u3(0), u3(1) keep their cx (order preserved); the other seven pair up into three 2-qubit gates + one left over
    assert kinds.count(_lib.GATE_DENSE2) == 3 and kinds.count(_lib.GATE_DENSE1) == 3 and kinds.count(_lib.GATE_CX) == 1
    st = FakeState(n, tile_qubits=9)
    st.init_basis0()
    st.apply_pass(passes[0].tile_qubits, passes[0].ops)
    ref = np.zeros(2 ** n, dtype=complex)
    ref[0] = 1.0
    for o in ops:
        mat = orc.cx_gate_matrix() if o.kind == _lib.GATE_CX else o.matrix
        ref = orc.apply_unitary(ref.reshape([2] * n), mat, list(o.qubits)).reshape(-1)
    assert np.max(np.abs(st.vec - ref)) < 1e-12


def test_swaps_are_eliminated_by_relabelling():
    """planner.eliminate_swaps: SWAP gates leave the stream (later gates renamed), the permutation they leave is at most two
    rounds of disjoint transpositions; ops + rounds = the original stream (oracle), swaps on the pinned low qubits stay
    gates, short streams are left alone, and a QFT's final swaps become one round."""
    from oracle import basicaer_oracle as orc
    from qiskit_sdk_py_b200 import _lib, circuits, lower, planner
    from qiskit_sdk_py_b200.lower import GateOp

    rng = np.random.RandomState(1)

    def ru(k):
        q, _ = np.linalg.qr(rng.standard_normal((2 ** k, 2 ** k)) + 1j * rng.standard_normal((2 ** k, 2 ** k)))
        return q

    def run(vec, ops, n):
        for o in ops:
            mat = np.diag(o.matrix) if o.kind in (_lib.GATE_DIAG1, _lib.GATE_DIAG2) else \
                orc.cx_gate_matrix() if o.kind == _lib.GATE_CX else o.matrix
            vec = orc.apply_unitary(vec.reshape([2] * n), mat, list(o.qubits)).reshape(-1)
        return vec

    n = 9
    swap = planner._SWAP
    for trial in range(20):
        ops = []
        for _ in range(25):
            a, b = (int(x) for x in rng.choice(n, 2, replace=False))
            r = rng.randint(4)
            ops.append(GateOp(_lib.GATE_DENSE2, [a, b], swap.copy()) if r < 2 else
                       GateOp(_lib.GATE_DENSE2, [a, b], ru(2)) if r == 2 else GateOp(_lib.GATE_DENSE1, [a], ru(1)))
        vec = rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)
        out, rounds = planner.eliminate_swaps(ops, n, low_qubits=2, min_swaps=1)
        assert 1 <= len(rounds) <= 2
        got = run(vec.copy(), out, n)
        for rnd in rounds:
            qs = [q for pair in rnd for q in pair]
            assert len(set(qs)) == len(qs) and min(qs) >= 2
            got = run(got, [GateOp(_lib.GATE_DENSE2, list(pair), swap) for pair in rnd], n)
        assert np.max(np.abs(got - run(vec.copy(), ops, n))) < 1e-12
        assert all(min(o.qubits) < 2 for o in out if planner.is_swap(o))
    few = [GateOp(_lib.GATE_DENSE2, [4, 5], swap.copy()), GateOp(_lib.GATE_DENSE1, [4], ru(1))]
    out, rounds = planner.eliminate_swaps(few, n)
    assert rounds == [] and out == few
    chain = [GateOp(_lib.GATE_DENSE2, [4 + i, 5 + i], swap.copy()) for i in range(4)]  # a 5-cycle: two launches for 4 swaps
    out, rounds = planner.eliminate_swaps(chain, n)
    assert rounds == [] and out == chain
    ins = circuits.qft(33)["experiments"][0]["instructions"]
    merged = planner.merge_ops([o for o in (lower.lower_gate(i["name"], i["qubits"], i.get("params")) for i in ins)
                                if o is not None])
    out, rounds = planner.eliminate_swaps(merged, 33)
    assert rounds == [[(q, 32 - q) for q in range(4, 16)]]
    assert len(planner.plan_passes(out, 33, tails=True)) == 6  # 10 with the swaps as gates


def test_in_tile_diagonals_behind_a_tail_join_the_pass():
    """plan_passes(tails=True): the controlled phases onto the pinned low qubits at the end of a QFT fan become ready
    only behind tail gates; they are appended to the pass's gates (diagonal gates commute with the all-diagonal tail)
    instead of costing a pass of their own.  Same state as the oracle, fewer passes than gates-then-tail planning."""
    from fake_state import FakeState
    from oracle import basicaer_oracle as orc
    from qiskit_sdk_py_b200 import _lib, circuits, lower, planner

    n = 12
    ins = circuits.qft(n, state_seed=5)["experiments"][0]["instructions"]
    ops = [o for o in (lower.lower_gate(i["name"], i["qubits"], i.get("params")) for i in ins if
                       i["name"] not in ("measure", "barrier")) if o is not None]
    merged = planner.merge_ops(ops)
    passes = planner.plan_passes(merged, n, tile_qubits=7, tails=True, pair_1q=False)
    assert sum(len(p.ops) + len(p.tail) for p in passes) == len(merged)
    passes = planner.plan_passes(merged, n, tile_qubits=7, tails=True)  # the default: state-preparation gates paired
    assert any(p.tail and any(o.kind == _lib.GATE_DIAG2 for o in p.ops) for p in passes)
    st = FakeState(n, tile_qubits=7)
    st.init_basis0()
    for p in passes:
        st.apply_pass(p.tile_qubits, p.ops, tail=p.tail)
    ref = np.zeros(2 ** n, dtype=complex)
    ref[0] = 1.0
    for o in ops:
        mat = np.diag(o.matrix) if o.kind in (_lib.GATE_DIAG1, _lib.GATE_DIAG2) else \
            orc.cx_gate_matrix() if o.kind == _lib.GATE_CX else o.matrix
        ref = orc.apply_unitary(ref.reshape([2] * n), mat, list(o.qubits)).reshape(-1)
    assert np.max(np.abs(st.vec - ref)) < 1e-12


def test_partial_plan_runs_what_must_run_and_leaves_a_successor_closed_rest():
    """plan_passes(must=..., leftover=...) (dist.ShardedState before a qubit exchange): every ``must`` gate and all its
    predecessors are planned, the rest is closed under successors, and planned + rest applied in that order give the
    oracle's state."""
    from fake_state import FakeState
    from oracle import basicaer_oracle as orc
    from qiskit_sdk_py_b200 import _lib, circuits, lower, planner

    n = 10
    ins = circuits.quantum_volume_instructions(n, 5, 11) + circuits.random_basis_instructions(n, 4, 3)
    ops = [o for o in (lower.lower_gate(i["name"], i["qubits"], i.get("params")) for i in ins) if o is not None]
    merged = planner.merge_ops(ops)
    victims = {7, 9}
    tainted, must = set(victims), set()
    for idx in range(len(merged) - 1, -1, -1):
        if set(merged[idx].qubits) & tainted:
            tainted |= set(merged[idx].qubits)
            must.add(idx)
    assert 0 < len(must) < len(merged)
    left = []
    passes = planner.plan_passes(merged, n, tile_qubits=6, tails=True, must=must, leftover=left)
    planned = sum(len(p.ops) + len(p.tail) for p in passes)
    assert planned + len(left) == len(merged) and left == sorted(left) and not (set(left) & must)
    for pos, idx in enumerate(left):       # closed under successors: nothing after a left gate on its qubits was planned
        for later in range(idx + 1, len(merged)):
            if set(merged[later].qubits) & set(merged[idx].qubits):
                assert later in left
    full = planner.plan_passes(merged, n, tile_qubits=6, tails=True)
    assert len(passes) <= len(full)
    st = FakeState(n, tile_qubits=6)
    st.init_basis0()
    for p in passes + planner.plan_passes([merged[i] for i in left], n, tile_qubits=6, tails=True):
        st.apply_pass(p.tile_qubits, p.ops, tail=p.tail)
    ref = np.zeros(2 ** n, dtype=complex)
    ref[0] = 1.0
    for o in ops:
        mat = np.diag(o.matrix) if o.kind in (_lib.GATE_DIAG1, _lib.GATE_DIAG2) else \
            orc.cx_gate_matrix() if o.kind == _lib.GATE_CX else o.matrix
        ref = orc.apply_unitary(ref.reshape([2] * n), mat, list(o.qubits)).reshape(-1)
    assert np.max(np.abs(st.vec - ref)) < 1e-12


@pytest.mark.parametrize("seed", range(12))
def test_full_host_pipeline_on_random_diagonal_heavy_circuits(seed):
    """merge_ops -> eliminate_swaps -> plan_passes(tails) with every planner rule on (diagonal tails, in-tile diagonals
    behind a tail, diagonal packing, paired 1-qubit gates, in-pass scheduling) -> the numpy test double pass by pass, then
    the swap rounds: the oracle's state for random circuits that are mostly diagonal gates, cx, swaps and rotations."""
    from fake_state import FakeState
    from oracle import basicaer_oracle as orc
    from qiskit_sdk_py_b200 import _lib, lower, planner

    rng = np.random.RandomState(500 + seed)
    n = int(rng.randint(7, 11))
    tile = int(rng.randint(5, min(n, 8) + 1))
    ins = []
    for _ in range(int(rng.randint(40, 120))):
        r = rng.randint(10)
        a, b = (int(x) for x in rng.choice(n, 2, replace=False))
        if r < 3:    # controlled phase in the simulator basis: u1, cx, u1, cx, u1
            lam = float(rng.uniform(-3, 3))
            ins += [{"name": "u1", "qubits": [a], "params": [lam / 2]}, {"name": "cx", "qubits": [a, b]},
                    {"name": "u1", "qubits": [b], "params": [-lam / 2]}, {"name": "cx", "qubits": [a, b]},
                    {"name": "u1", "qubits": [b], "params": [lam / 2]}]
        elif r < 5:
            ins.append({"name": str(rng.choice(["u1", "rz"])), "qubits": [a], "params": [float(rng.uniform(-3, 3))]})
        elif r < 6:
            ins += [{"name": "cx", "qubits": [a, b]}, {"name": "cx", "qubits": [b, a]}, {"name": "cx", "qubits": [a, b]}]
        elif r < 8:
            ins.append({"name": "u3", "qubits": [a], "params": [float(x) for x in rng.uniform(-3, 3, 3)]})
        elif r < 9:
            ins.append({"name": "cx", "qubits": [a, b]})
        else:
            ins.append({"name": "u2", "qubits": [a], "params": [float(x) for x in rng.uniform(-3, 3, 2)]})
    ops = [o for o in (lower.lower_gate(i["name"], i["qubits"], i.get("params")) for i in ins) if o is not None]
    merged = planner.merge_ops(ops)
    out, rounds = planner.eliminate_swaps(merged, n, low_qubits=min(planner.LOW_QUBITS, tile - 2), min_swaps=2)
    passes = planner.plan_passes(out, n, tile_qubits=tile, tails=True)
    st = FakeState(n, tile_qubits=tile)
    vec = rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)
    st.vec = vec.copy()
    for p in passes:
        st.apply_pass(p.tile_qubits, p.ops, tail=p.tail)
    got = st.vec
    for rnd in rounds:
        for pair in rnd:
            got = orc.apply_unitary(got.reshape([2] * n), planner._SWAP, list(pair)).reshape(-1)
    ref = vec.copy()
    for o in ops:
        mat = np.diag(o.matrix) if o.kind in (_lib.GATE_DIAG1, _lib.GATE_DIAG2) else \
            orc.cx_gate_matrix() if o.kind == _lib.GATE_CX else o.matrix
        ref = orc.apply_unitary(ref.reshape([2] * n), mat, list(o.qubits)).reshape(-1)
    assert np.max(np.abs(got - ref)) < 1e-10 * max(1.0, np.max(np.abs(ref)))


def test_device_statevector_handle_against_oracle():
    """statevector.DeviceStatevector (here over the numpy test double): expectation values incl. qargs, linear
    combinations and signed labels, probabilities in the caller's qubit order, seeded sampling -- against the oracle's
    restatement of Statevector._expectation_value_pauli (statevector.py:377-408) and plain numpy."""
    from fake_state import FakeState
    from oracle import basicaer_oracle as orc
    from qiskit_sdk_py_b200 import circuits
    from qiskit_sdk_py_b200.statevector import DeviceStatevector

    n = 6
    ins = circuits.random_basis_instructions(n, 6, 5)
    sv = DeviceStatevector.from_instructions(n, ins, state_factory=lambda k: FakeState(k, tile_qubits=4))
    vec = sv.to_numpy()
    assert sv.num_qubits == n and sv.dim == 64 and abs(sv.norm() - 1.0) < 1e-12
    for label in ("ZIIXYZ", "XXXXXX", "IIIIIZ", "YIZIXI"):
        assert abs(sv.expectation_value(label) - orc.expval_pauli(vec, label)) < 1e-12
    assert abs(sv.expectation_value("-iZIIXYZ") + 1j * orc.expval_pauli(vec, "ZIIXYZ")) < 1e-12
    # qargs: "XZ" on qubits [1, 4] = X on qubit 4, Z on qubit 1
    assert abs(sv.expectation_value("XZ", qargs=[1, 4]) - orc.expval_pauli(vec, "IXIIZI")) < 1e-12
    combo = [("ZIIIII", 0.5), ("IXIIII", -1.5j)]
    want = 0.5 * orc.expval_pauli(vec, "ZIIIII") - 1.5j * orc.expval_pauli(vec, "IXIIII")
    assert abs(sv.expectation_value(combo) - want) < 1e-12
    p = np.abs(vec) ** 2
    idx = np.arange(64)
    got = sv.probabilities([4, 1])
    want = np.zeros(4)
    for i in range(64):
        want[((i >> 4) & 1) | (((i >> 1) & 1) << 1)] += p[i]
    assert np.max(np.abs(got - want)) < 1e-13
    assert np.max(np.abs(sv.probabilities() - p)) < 1e-13
    mem = sv.sample_memory(200, seed=3)
    cdf = p.cumsum()
    exp = (cdf / cdf[-1]).searchsorted(np.random.RandomState(3).random_sample(200), side="right")
    assert mem == [format(int(s), "06b") for s in exp]
    assert sum(sv.sample_counts(500, qargs=[0, 5], seed=1).values()) == 500
    sv.evolve([circuits.u3(0.3, 0.2, 0.1, 2)])
    assert abs(sv.norm() - 1.0) < 1e-12


@pytest.mark.parametrize("name", ["midcircuit_8", "teleport", "reset_midcircuit", "if_statement",
                                  "transpiled_init_reset_cond_counts"])
def test_shot_batched_execution_equals_shot_by_shot(name):
    """circuits that cannot be sampled run 2^b shots side by side in one (n + b)-qubit vector
    (per-shot outcomes from the same RandomState stream, per-shot collapse, masked conditional gates) -- memory and
    counts identical to the shot-by-shot path and to the reference's recorded output (qasm_simulator.py:512-614)."""
    case = golden_io.load_case(name)
    batched = simulator.QasmSimulatorB200(state_factory=FakeState, max_qubits=24)
    res_b = batched.run(copy.deepcopy(case["qobj"]), **case.get("options", {}))
    assert batched.last_stats["shot_batch_bits"] >= 3
    serial = simulator.QasmSimulatorB200(state_factory=FakeState, max_qubits=24)
    serial.MIN_BATCH_BITS = 99
    res_s = serial.run(copy.deepcopy(case["qobj"]), **case.get("options", {}))
    assert serial.last_stats["shot_batch_bits"] == 0
    assert res_b["results"][0]["data"] == res_s["results"][0]["data"]
    helpers.compare_result(res_b, case)
