"""Multi-rank path on CPU: world_size 2 and 4 over gloo, the shard replaced by the numpy test double.
Exercises the sharding logic of dist.ShardedState -- localisation of gates on global qubits, qubit
swaps (pack / exchange / unpack), layout restore, distributed P(qubit), collapse and sampling --
against the single-process oracle."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, seed, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fake_state import FakeState
    from oracle import basicaer_oracle as orc
    from qiskit_sdk_py_b200 import _lib, circuits, lower
    from qiskit_sdk_py_b200.dist import ShardedState

    ins = circuits.random_basis_instructions(n, 10, seed) + circuits.quantum_volume_instructions(n, 3, seed)
    ops = [lower.lower_gate(i["name"], i["qubits"], i.get("params")) for i in ins]
    ops = [o for o in ops if o is not None]
    st = ShardedState(n, lambda k: FakeState(k, tile_qubits=4), chunk_amps=8)
    st.init_basis0(np.exp(0.2j))
    st.apply_ops(ops)
    # oracle
    ref = np.zeros(2 ** n, dtype=complex)
    ref[0] = np.exp(0.2j)
    for o in ops:
        if o.kind in (_lib.GATE_DIAG1, _lib.GATE_DIAG2):
            mat = np.diag(o.matrix)
        elif o.kind == _lib.GATE_CX:
            mat = orc.cx_gate_matrix()
        else:
            mat = o.matrix
        ref = orc.apply_unitary(ref.reshape([2] * n), mat, list(o.qubits)).reshape(-1)
    checks = {"swaps": st.swaps, "fused": st.fused_exchanges}
    # distributed probabilities on a permuted layout
    p = np.abs(ref) ** 2
    idx = np.arange(2 ** n)
    for q in range(n):
        p0, p1 = st.prob_qubit(q)
        mask = ((idx >> q) & 1).astype(bool)
        assert abs(p0 - p[~mask].sum()) < 1e-12 and abs(p1 - p[mask].sum()) < 1e-12, (q, p0, p1)
    assert abs(st.norm2() - 1.0) < 1e-12
    # sampling from a permuted layout: sample_all restores the layout and chains numpy's running sum through the
    # ranks -- the indices are exactly searchsorted(cumsum(|psi|^2) / total, u, 'right') over the WHOLE register
    # (qasm_simulator.py:213), with psi the sharded state's own amplitudes
    u = np.random.RandomState(7).random_sample(4000)
    samples, total = st.sample_all(u)
    assert st.pos == list(range(n))
    full_now = st.gather_statevector()
    cdf = (np.abs(full_now) ** 2).cumsum()
    assert total == cdf[-1]
    assert np.array_equal(samples, (cdf / cdf[-1]).searchsorted(u, side="right"))
    hist = np.bincount(samples, minlength=2 ** n) / 4000.0
    assert np.abs(hist - p).max() < 0.05
    st.apply_ops(ops[:6])  # leave the identity layout again for the checks below
    for o in ops[:6]:
        if o.kind in (_lib.GATE_DIAG1, _lib.GATE_DIAG2):
            mat = np.diag(o.matrix)
        elif o.kind == _lib.GATE_CX:
            mat = orc.cx_gate_matrix()
        else:
            mat = o.matrix
        ref = orc.apply_unitary(ref.reshape([2] * n), mat, list(o.qubits)).reshape(-1)
    p = np.abs(ref) ** 2
    # Pauli expectation values on the permuted layout: X / Y on global qubits are swapped in, Z on a global
    # qubit is the sign of the rank bit (reference: quantum_info/states/statevector.py:377-408 via the oracle)
    prng = np.random.RandomState(100 + n)
    labels = ["I" * n, "Z" * n, "X" + "I" * (n - 1), "Y" + "Z" * (n - 1), "I" * (n - 1) + "X"]
    labels += ["".join(prng.choice(list("IXYZ"), n)) for _ in range(6)]
    for label in labels:
        if sum(c in "XY" for c in label) > st.n_local:
            continue
        want = orc.expval_pauli(ref, label)
        got = st.expval_pauli(label)
        assert abs(got - want) < 1e-12, (label, got, want)
    # collapse a (possibly global) qubit
    qm = n - 1
    probs = st.prob_qubit(qm)
    st.collapse(qm, 1, probs[1], is_reset=False)
    mask = ((idx >> qm) & 1).astype(bool)
    ref2 = np.where(mask, ref, 0) / np.sqrt(probs[1])
    full = st.gather_statevector()
    assert st.pos == list(range(n))
    err = float(np.max(np.abs(full - ref2)))
    assert err < 1e-12, err
    checks["err"] = err
    if rank == 0:
        np.save(os.path.join(out_dir, "ok_%d_%d.npy" % (world, n)), np.array([checks["swaps"], err, checks["fused"]]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 5), (2, 7), (4, 6), (4, 8)])
def test_sharded_state_matches_oracle(tmp_path, world, n):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, 11 + n, str(tmp_path)), nprocs=world, join=True)
    res = np.load(os.path.join(str(tmp_path), "ok_%d_%d.npy" % (world, n)))
    assert res[0] > 0, "the test circuit should have needed qubit swaps"
    assert res[1] < 1e-12
    if world == 4:
        assert res[2] > 0, "world 4 should have used the fused multi-qubit exchange"


# ---- the BasicAer mirror simulators running SPMD on the sharded state ---------------------------------
# Golden fixtures (outputs of the unmodified reference) whose circuits have enough qubits for `world` ranks:
# seeded counts / memory must be IDENTICAL, amplitudes within 1e-12, on every rank.
SPMD_CASES = {
    2: ["if_statement", "memory_two_registers", "multi_registers_counts", "multi_registers_state",
        "qasm_initial_statevector", "reset_midcircuit", "sampler_marginal_subset", "sampler_partial_qubit",
        "sampler_single_qubit", "sv_global_phase", "sv_initial_statevector", "sv_reset_midcircuit", "teleport",
        "transpiled_init_reset_cond_counts", "transpiled_grover_5_counts", "qv_5_state", "random_6_state",
        "qft_5", "transpiled_lib_gates_l1_state"],
    4: ["memory_two_registers", "multi_registers_counts", "sampler_marginal_subset", "sampler_partial_qubit",
        "sampler_single_qubit", "sv_global_phase", "transpiled_init_reset_cond_counts", "qv_8_state",
        "random_9_state", "transpiled_qft_lib_9_state", "transpiled_lib_gates_l0_state"],
}
SPMD_UNITARY = ["random_2_unitary", "random_4_unitary", "unitary_five_circuits", "unitary_global_phase_x4",
                "unitary_initial_unitary", "unitary_random", "multi_registers_unitary"]
# all qubits measured: the exact-order CDF is chained through the ranks (ShardedState.sample_all), counts identical
SPMD_FULL_MEASURE = {2: ["qv_5_counts", "random_6_counts", "transpiled_lib_gates_l0_counts"],
                     4: ["qv_8_counts", "random_9_counts", "transpiled_qv_lib_10_counts"]}


def _spmd_worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import golden_io
    import helpers
    from fake_state import FakeState
    from qiskit_sdk_py_b200.dist import distributed_simulator

    def factory(backend):
        return distributed_simulator(backend, shard_factory=lambda k: FakeState(k, tile_qubits=4), chunk_amps=8,
                                     max_qubits=24)

    done = 0
    for name in SPMD_CASES[world] + SPMD_FULL_MEASURE[world] + SPMD_UNITARY:
        case = golden_io.load_case(name)
        result = helpers.run_case_with(factory, case)
        helpers.compare_result(result, case, amp_tol=1e-12)
        done += 1
    # a seed the backend draws itself is rank 0's on every rank
    from qiskit_sdk_py_b200 import circuits

    qobj = circuits.quantum_volume(6, shots=64)
    qobj["config"].pop("seed_simulator", None)
    res = factory("qasm_simulator").run(qobj)
    seeds = [None] * world
    dist.all_gather_object(seeds, (res["results"][0]["seed_simulator"], res["results"][0]["data"]["counts"]))
    assert all(s == seeds[0] for s in seeds), seeds
    # the call sequence of bench.py's N > 1 end-to-end leg (dist.bench_main), at a size the oracle can check
    import copy as _copy

    from oracle import basicaer_oracle as _orc
    from qiskit_sdk_py_b200.dist import DEFAULT_CHUNK_AMPS

    nq = 7
    n_loc = nq - int(np.log2(world))
    qobj = circuits.quantum_volume(nq, nq, 1234, shots=256, seed_simulator=42)
    sim = distributed_simulator("qasm_simulator", shard_factory=lambda k: FakeState(k, tile_qubits=4),
                                chunk_amps=min(DEFAULT_CHUNK_AMPS, 1 << (n_loc - 1)), max_qubits=nq)
    sim.run(_copy.deepcopy(qobj))
    res = sim.run(_copy.deepcopy(qobj))
    counts = res["results"][0]["data"]["counts"]
    assert sum(counts.values()) == 256
    assert counts == _orc.OracleQasmSimulator().run(_copy.deepcopy(qobj))["results"][0]["data"]["counts"]
    # a 3-qubit unitary on global qubits of a sharded state (swap-in of its qubits, then the dense-k path)
    from oracle import basicaer_oracle as orc

    rng = np.random.RandomState(3)
    q, r = np.linalg.qr(rng.standard_normal((8, 8)) + 1j * rng.standard_normal((8, 8)))
    u3q = q * (np.diag(r) / np.abs(np.diag(r)))
    n = 6
    qobj = {"qobj_id": "k3", "type": "QASM", "schema_version": "1.3.0", "header": {},
            "config": {"shots": 1, "memory_slots": 0, "n_qubits": n},
            "experiments": [{"header": {"name": "k3", "memory_slots": 0, "n_qubits": n, "creg_sizes": [], "clbit_labels": [],
                                        "qreg_sizes": [["q", n]], "qubit_labels": [["q", i] for i in range(n)],
                                        "global_phase": 0.0},
                             "config": {"n_qubits": n, "memory_slots": 0},
                             "instructions": [{"name": "u2", "qubits": [5], "params": [0.0, np.pi]},
                                              {"name": "cx", "qubits": [5, 0]},
                                              {"name": "unitary", "qubits": [5, 1, 4], "params": [u3q]},
                                              {"name": "u3", "qubits": [4], "params": [0.3, 0.2, 0.1]}]}]}
    import copy

    got = factory("statevector_simulator").run(copy.deepcopy(qobj))["results"][0]["data"]["statevector"]
    ref = orc.OracleQasmSimulator(show_final_state=True).run(copy.deepcopy(qobj))["results"][0]["data"]["statevector"]
    assert np.max(np.abs(got - ref)) < 1e-12
    if rank == 0:
        np.save(os.path.join(out_dir, "spmd_%d.npy" % world), np.array([done]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_simulators_run_spmd_on_sharded_state(tmp_path, world):
    """dist.distributed_simulator: QasmSimulatorB200 / StatevectorSimulatorB200 / UnitarySimulatorB200 with the
    state sharded over `world` gloo ranks reproduce the reference's golden outputs (mid-circuit measure, reset, conditionals,
    partial-register sampling, initial_statevector, global phase, memory) on every rank."""
    port = _free_port()
    mp.spawn(_spmd_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = np.load(os.path.join(str(tmp_path), "spmd_%d.npy" % world))
    assert res[0] == len(SPMD_CASES[world]) + len(SPMD_FULL_MEASURE[world]) + len(SPMD_UNITARY)


def _provider_worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle", "ref_env")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import warnings

    import ref_import

    ref_import.import_reference()
    warnings.simplefilter("ignore")
    import qiskit
    from fake_state import FakeState
    from qiskit_sdk_py_b200.provider import B200Provider

    provider = B200Provider(simulator_kwargs={"state_factory": FakeState, "max_qubits": 12}, distributed=True,
                            distributed_kwargs={"shard_factory": lambda k: FakeState(k, tile_qubits=4),
                                                "chunk_amps": 8, "max_qubits": 24})
    qc = qiskit.QuantumCircuit(4, 4)
    qc.h(3)
    qc.cx(3, 0)
    qc.ry(0.7, 2)
    qc.cx(2, 1)
    qc.measure([0, 1, 2, 3], [0, 1, 2, 3])
    ours = qiskit.execute(qc, provider.get_backend("qasm_simulator"), shots=500, seed_simulator=9, memory=True).result()
    ref = qiskit.execute(qc, qiskit.BasicAer.get_backend("qasm_simulator"), shots=500, seed_simulator=9,
                         memory=True).result()
    assert ours.get_counts() == ref.get_counts() and ours.get_memory() == ref.get_memory()
    qc2 = qiskit.QuantumCircuit(5)
    qc2.h(4)
    qc2.cx(4, 1)
    qc2.rz(0.3, 4)
    qc2.ccx(4, 3, 0)
    sv = qiskit.execute(qc2, provider.get_backend("statevector_simulator")).result().get_statevector()
    sv_ref = qiskit.execute(qc2, qiskit.BasicAer.get_backend("statevector_simulator")).result().get_statevector()
    assert np.max(np.abs(np.asarray(sv) - np.asarray(sv_ref))) < 1e-12
    assert provider.get_backend("qasm_simulator").configuration().n_qubits == 24
    qc3 = qiskit.QuantumCircuit(3)
    qc3.h(2)
    qc3.cx(2, 0)
    qc3.t(1)
    un = qiskit.execute(qc3, provider.get_backend("unitary_simulator")).result().get_unitary()
    un_ref = qiskit.execute(qc3, qiskit.BasicAer.get_backend("unitary_simulator")).result().get_unitary()
    assert np.max(np.abs(np.asarray(un) - np.asarray(un_ref))) < 1e-12
    if rank == 0:
        np.save(os.path.join(out_dir, "provider_%d.npy" % world), np.array([1]))
    dist.destroy_process_group()


@pytest.mark.skipif(not os.path.isdir("/root/reference/qiskit"), reason="reference/qiskit not present")
def test_distributed_provider_with_reference_execute(tmp_path):
    """B200Provider(distributed=True) under the reference's own execute(): 2 gloo ranks, same seeded counts,
    memory and statevector as BasicAer on every rank."""
    port = _free_port()
    mp.spawn(_provider_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(os.path.join(str(tmp_path), "provider_2.npy"))


def _subgroup_worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import golden_io
    import helpers
    from fake_state import FakeState
    from qiskit_sdk_py_b200.dist import distributed_simulator

    group = dist.new_group(ranks=[2, 3])  # every rank must call new_group
    if rank in (2, 3):
        def factory(backend):
            return distributed_simulator(backend, shard_factory=lambda k: FakeState(k, tile_qubits=4), chunk_amps=8,
                                         max_qubits=24, group=group)

        for name in ["qv_5_counts", "random_6_state", "sampler_marginal_subset", "teleport", "random_4_unitary"]:
            case = golden_io.load_case(name)
            helpers.compare_result(helpers.run_case_with(factory, case), case, amp_tol=1e-12)
        np.save(os.path.join(out_dir, "sub_%d.npy" % rank), np.array([1]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_state_on_a_strict_subgroup(tmp_path):
    """exchanges address their peers by GLOBAL rank (torch's P2POp semantics) -- a state sharded over
    ranks [2, 3] of a 4-rank job must not talk to ranks 0 / 1."""
    port = _free_port()
    mp.spawn(_subgroup_worker, args=(4, port, str(tmp_path)), nprocs=4, join=True)
    assert os.path.exists(os.path.join(str(tmp_path), "sub_2.npy")) and os.path.exists(os.path.join(str(tmp_path), "sub_3.npy"))


def _parity_gate_worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fake_state import FakeState
    from qiskit_sdk_py_b200.dist import parity_gate

    kwargs = {"shard_factory": lambda k: FakeState(k, tile_qubits=4), "chunk_amps": 8, "max_qubits": 24}
    cases = ["midcircuit_8", "sv_midcircuit_10", "random_6_unitary", "sampler_marginal_subset", "qv_11_counts",
             "transpiled_init_reset_cond_counts", "bell_seed42"]
    summary, details = parity_gate(kwargs, world, cases=cases, device="cpu")
    fits = [c for c in cases if c not in ("bell_seed42",) and not (world == 4 and c == "transpiled_init_reset_cond_counts"
                                                                    and False)]
    assert summary.startswith("golden %d/%d identical" % (len(details), len(details))) and "FAILED" not in summary, summary
    assert "bell_seed42" not in details          # 2 qubits do not leave two local qubits per rank
    assert "midcircuit_8" in details and details["midcircuit_8"]["ok"]
    assert ("sampler_marginal_subset" in details) and details["sampler_marginal_subset"]["ok"]  # 2 local qubits at world 8
    if rank == 0:
        np.save(os.path.join(out_dir, "gate_%d.npy" % world), np.array([len(details), len(fits)]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_bench_parity_gate(tmp_path, world):
    """The golden parity gate bench.py --gpus N runs before its timed region (dist.parity_gate), here over gloo."""
    port = _free_port()
    mp.spawn(_parity_gate_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = np.load(os.path.join(str(tmp_path), "gate_%d.npy" % world))
    assert res[0] >= 5


def _swap_relabel_worker(rank, world, port, n, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fake_state import FakeState
    from oracle import basicaer_oracle as orc
    from qiskit_sdk_py_b200 import _lib, circuits, lower
    from qiskit_sdk_py_b200.dist import ShardedState

    # a QFT with its final swaps (qubit n-1 <-> 0, ...: the home qubits of the global bits end up in the LOWEST local
    # bits), then a few more gates and swaps on global qubits behind deferred gates
    ins = [i for i in circuits.qft(n, state_seed=3)["experiments"][0]["instructions"] if i["name"] not in ("measure", "barrier")]
    ins += circuits.random_basis_instructions(n, 3, 21)
    for a, b in ((n - 1, 2), (n - 2, n - 1), (0, n - 1)):
        ins += [{"name": "cx", "qubits": [a, b]}, {"name": "cx", "qubits": [b, a]}, {"name": "cx", "qubits": [a, b]}]
        ins += circuits.random_basis_instructions(n, 1, 30 + a)
    ops = [o for o in (lower.lower_gate(i["name"], i["qubits"], i.get("params")) for i in ins) if o is not None]
    st = ShardedState(n, lambda k: FakeState(k, tile_qubits=5), chunk_amps=16)
    st.init_basis0()
    st.apply_ops(ops)
    assert st.relabelled_swaps >= n // 2 + 3
    swaps_before_restore = st.swaps
    full = st.gather_statevector()  # restores the identity layout
    assert st.pos == list(range(n))
    ref = np.zeros(2 ** n, dtype=complex)
    ref[0] = 1.0
    for o in ops:
        mat = np.diag(o.matrix) if o.kind in (_lib.GATE_DIAG1, _lib.GATE_DIAG2) else \
            orc.cx_gate_matrix() if o.kind == _lib.GATE_CX else o.matrix
        ref = orc.apply_unitary(ref.reshape([2] * n), mat, list(o.qubits)).reshape(-1)
    assert np.max(np.abs(full - ref)) < 1e-12
    if rank == 0:
        np.save(os.path.join(out_dir, "relabel_%d.npy" % world), np.array([st.relabelled_swaps, swaps_before_restore, st.swaps]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 10), (4, 11), (8, 12)])
def test_swaps_are_relabelled_on_the_sharded_state(tmp_path, world, n):
    """ShardedState.apply_ops turns cx a,b; cx b,a; cx a,b into a relabelling of the two logical qubits (no exchange, no
    pass), also for global qubits and behind deferred gates; restore_identity_layout settles the layout once and keeps the
    lowest local bits out of the exchange.  Same state as the oracle applying every cx."""
    port = _free_port()
    mp.spawn(_swap_relabel_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    res = np.load(os.path.join(str(tmp_path), "relabel_%d.npy" % world))
    assert res[0] >= n // 2 + 3


def _overlap_worker(rank, world, port, n, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fake_state import FakeState
    from oracle import basicaer_oracle as orc
    from qiskit_sdk_py_b200 import _lib, circuits, lower
    from qiskit_sdk_py_b200.dist import ShardedState

    ins = circuits.quantum_volume_instructions(n, 6, 5) + circuits.random_basis_instructions(n, 6, 9)
    ops = [o for o in (lower.lower_gate(i["name"], i["qubits"], i.get("params")) for i in ins) if o is not None]
    results = []
    for overlap, carry in ((True, False), (False, False), (True, True), (False, True)):
        st = ShardedState(n, lambda k: FakeState(k, tile_qubits=5), chunk_amps=16, overlap=overlap)
        st.carry_over = carry  # gates an exchange does not need wait for the passes after it (plan_passes(must=...))
        st.MIN_OVERLAP_GATES = 2
        st.MIN_OVERLAP_SWAPS = 1
        st.init_basis0()
        st.apply_ops(ops)
        results.append((st.gather_statevector(), st.overlapped_exchanges, st.swaps, st.shard.passes_launched))
    ref = np.zeros(2 ** n, dtype=complex)
    ref[0] = 1.0
    for o in ops:
        mat = np.diag(o.matrix) if o.kind in (_lib.GATE_DIAG1, _lib.GATE_DIAG2) else \
            orc.cx_gate_matrix() if o.kind == _lib.GATE_CX else o.matrix
        ref = orc.apply_unitary(ref.reshape([2] * n), mat, list(o.qubits)).reshape(-1)
    assert np.max(np.abs(results[0][0] - ref)) < 1e-12
    assert np.max(np.abs(results[1][0] - ref)) < 1e-12
    assert np.max(np.abs(results[2][0] - ref)) < 1e-12
    assert np.max(np.abs(results[3][0] - ref)) < 1e-12
    assert results[0][1] > 0, "the circuit should have had exchanges with late passes to overlap"
    assert results[1][1] == 0 and results[3][1] == 0
    assert results[3][3] <= results[1][3], "carry-over must not cost passes"
    if rank == 0:
        np.save(os.path.join(out_dir, "ovl_%d.npy" % world), np.array([results[0][1], results[0][2]]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 11), (4, 12)])
def test_exchange_overlapped_with_commuting_passes(tmp_path, world, n):
    """ShardedState with overlap=True holds back the locally runnable gates that commute with the coming exchange
    (no successor-or-self on a swapped bit) and applies them sub-block by sub-block between the partial exchanges
    (qsv_apply_pass_sub + one partner per step): same state as the oracle and as the plain schedule."""
    port = _free_port()
    mp.spawn(_overlap_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    res = np.load(os.path.join(str(tmp_path), "ovl_%d.npy" % world))
    assert res[0] > 0
