"""End-to-end GPU parity: the product simulators (CUDA path through the C-ABI) against the outputs
the unmodified reference produced for the same Qobj (tests/golden/, see make_golden.py).
Tolerance (north_star): max-abs amplitude error <= 1e-10 in complex128; counts / memory identical
for identical seeds."""
import copy

import numpy as np
import pytest

import golden_io
import helpers
from qiskit_sdk_py_b200 import _lib, circuits, simulator

pytestmark = pytest.mark.gpu

ALL = golden_io.list_cases()


def gpu_factory(backend):
    return simulator.get_simulator(backend)


@pytest.mark.parametrize("name", ALL)
def test_cuda_matches_reference_golden(name):
    case = golden_io.load_case(name)
    before = _lib.launch_count()
    result = helpers.run_case_with(gpu_factory, case)
    assert _lib.launch_count() > before, "no libqsv kernel was launched"
    helpers.compare_result(result, case, amp_tol=helpers.AMP_TOL)


def test_bell_config1():
    """BASELINE config 1: {'00': 511, '11': 513} for seed_simulator=42."""
    res = gpu_factory("qasm_simulator").run(circuits.bell())
    assert res["results"][0]["data"]["counts"] == {"0x0": 511, "0x3": 513}


def _fits(n):
    import gc

    import torch

    gc.collect()
    torch.cuda.empty_cache()
    free, _total = torch.cuda.mem_get_info()
    return (16 << n) + (6 << 30) < free


@pytest.mark.parametrize("n", [24, 30, 33])
def test_qft_known_answer_at_full_size(n):
    """A known-answer test that scales: the QFT circuit of BASELINE config 2's family (basis gates, final swaps)
    maps |x> to the DFT column  amp[y] = exp(2 pi i x y / 2^n) / 2^(n/2)  -- checked here up to the 33-qubit
    register of BASELINE config 4 (128 GiB, ~2700 instructions) on 512 sampled amplitudes, plus the norm."""
    import torch

    from qiskit_sdk_py_b200 import lower
    from qiskit_sdk_py_b200.engine import DeviceState

    if not _fits(n):
        pytest.skip("not enough device memory for %d qubits" % n)
    rng = np.random.RandomState(n)
    x = int(rng.randint(1, 2 ** 31)) | (1 << (n - 1)) | 1
    x &= (1 << n) - 1
    ins = [{"name": "x", "qubits": [q]} for q in range(n) if (x >> q) & 1]
    ins += circuits.qft(n)["experiments"][0]["instructions"]
    ops = [o for o in (lower.lower_gate(i["name"], i["qubits"], i.get("params")) for i in ins) if o is not None]
    st = DeviceState(n)
    st.init_basis0()
    st.apply_ops(ops)
    assert st.passes_launched <= 20
    assert abs(st.norm2() - 1.0) < 1e-10
    ys = [0, 1, 2, (1 << n) - 1, 1 << (n - 1)] + [int(v) for v in rng.randint(0, 2 ** 31, 507) * 5 % (1 << n)]
    idx = torch.tensor([2 * y for y in ys] + [2 * y + 1 for y in ys], dtype=torch.int64, device=st.state.device)
    vals = st.state.index_select(0, idx).cpu().numpy()
    got = vals[: len(ys)] + 1j * vals[len(ys):]
    frac = np.array([((x * y) % (1 << n)) / float(1 << n) for y in ys])  # exact integer arithmetic for the phase
    want = np.exp(2j * np.pi * frac) / np.sqrt(float(1 << n))
    assert np.max(np.abs(got - want)) * np.sqrt(float(1 << n)) < 1e-9   # relative to the amplitude modulus
    assert np.max(np.abs(got - want)) < helpers.AMP_TOL


@pytest.mark.parametrize("n", [26, 28, 33])
def test_large_properties(n):
    """Sizes beyond the reference's 24-qubit cap: norm preservation and the inverse-circuit echo
    (SURVEY 8(d) config 4 correctness plan)."""
    from qiskit_sdk_py_b200.engine import DeviceState
    from qiskit_sdk_py_b200 import lower

    if not _fits(n):
        pytest.skip("not enough device memory for %d qubits" % n)
    ins = circuits.quantum_volume_instructions(n, 4, seed=99)
    ops = [lower.lower_gate(i["name"], i["qubits"], i.get("params")) for i in ins]
    st = DeviceState(n)
    st.init_basis0()
    st.apply_ops(ops)
    assert abs(st.norm2() - 1.0) < 1e-10
    p0, p1 = st.prob_qubit(n - 1)
    assert abs(p0 + p1 - 1.0) < 1e-10 and 0.0 < p0 < 1.0
    inv = [lower.matrix_op(np.conj(op.matrix.T), op.qubits) for op in reversed(ops)]
    st.apply_ops(inv)
    head = st.state[:2].cpu().numpy()
    assert abs(complex(head[0], head[1]) - 1.0) < 1e-10
    u = np.random.RandomState(1).random_sample(64)
    got, total = st.sample(list(range(n)), u, exact_order=False)
    assert np.all(got == 0) and abs(total - 1.0) < 1e-10


def test_no_fallback_when_library_missing(monkeypatch):
    """The product path must fail loudly without the CUDA library."""
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libqsv.so")
    with pytest.raises(_lib.QsvLibraryError):
        gpu_factory("qasm_simulator").run(circuits.bell())


def _oracle_factory(backend):
    from oracle import basicaer_oracle as orc

    if backend == "qasm_simulator":
        return orc.OracleQasmSimulator()
    if backend == "statevector_simulator":
        return orc.OracleQasmSimulator(show_final_state=True)
    return orc.OracleUnitarySimulator()


def _same_as_oracle(backend, qobj, **options):
    got = gpu_factory(backend).run(copy.deepcopy(qobj), **options)
    ref = _oracle_factory(backend).run(copy.deepcopy(qobj), **options)
    for g, r in zip(got["results"], ref["results"]):
        assert g["shots"] == r["shots"]
        if "seed_simulator" in qobj["config"]:
            assert g["seed_simulator"] == r["seed_simulator"]
        for key in ("counts", "memory"):
            assert (key in g["data"]) == (key in r["data"])
            if key in r["data"]:
                assert g["data"][key] == r["data"][key]
                if key == "counts":
                    assert list(g["data"][key]) == list(r["data"][key])
        for key in ("statevector", "unitary"):
            assert (key in g["data"]) == (key in r["data"])
            if key in r["data"]:
                assert np.max(np.abs(np.asarray(g["data"][key]) - r["data"][key])) <= helpers.AMP_TOL
    return got


def test_edge_max_shots_and_memory():
    """max_shots = 65536 (qasm_simulator.py:71) with memory=True."""
    q = circuits.bell(shots=65536, seed_simulator=9)
    q["config"]["memory"] = True
    got = _same_as_oracle("qasm_simulator", q)
    assert len(got["results"][0]["data"]["memory"]) == 65536


def test_edge_wide_classical_register():
    """More than 62 memory slots: classical memory exceeds int64, Python-int path."""
    ins = [circuits.u2(0.0, np.pi, 0), circuits.cx(0, 1), circuits.u3(1.0, 0.2, 0.3, 2),
           circuits.measure(0, 69), circuits.measure(1, 3), circuits.measure(2, 64), circuits.measure(0, 0)]
    exp = circuits.make_experiment("wide", 3, ins, memory_slots=70)
    _same_as_oracle("qasm_simulator", circuits.make_qobj(exp, shots=300, seed_simulator=4, memory=True))


def test_edge_empty_and_single_qubit_circuits():
    _same_as_oracle("statevector_simulator", circuits.make_qobj(circuits.make_experiment("empty", 3, []), shots=1))
    _same_as_oracle("qasm_simulator", circuits.make_qobj(
        circuits.make_experiment("only_measure", 2, circuits.measure_all(2), memory_slots=2), shots=20, seed_simulator=1))
    ins = [circuits.u3(0.7, 0.1, 0.2, 0), {"name": "sx", "qubits": [0]}, circuits.u1(0.3, 0), circuits.measure(0, 0)]
    _same_as_oracle("qasm_simulator", circuits.make_qobj(
        circuits.make_experiment("one", 1, ins, memory_slots=1), shots=500, seed_simulator=3, memory=True))
    _same_as_oracle("statevector_simulator", circuits.make_qobj(circuits.make_experiment("one_sv", 1, ins[:3]), shots=1))
    _same_as_oracle("unitary_simulator", circuits.make_qobj(circuits.make_experiment("one_u", 1, ins[:3]), shots=1))


@pytest.mark.parametrize("n", [7, 9])
def test_unitary_simulator_larger(n):
    ins = circuits.random_basis_instructions(n, 6, 100 + n)
    _same_as_oracle("unitary_simulator", circuits.make_qobj(circuits.make_experiment("u%d" % n, n, ins), shots=1))


def test_midcircuit_measure_reset_conditional_large_shots():
    """Non-sampled path (per-shot simulation with the prefix snapshot): RNG order must match the reference."""
    ins = [circuits.u2(0.0, np.pi, 0), circuits.cx(0, 1), circuits.u3(0.9, 0.1, 0.4, 2), circuits.cx(2, 3),
           {"name": "measure", "qubits": [0], "memory": [0], "register": [0]},
           {"name": "bfunc", "mask": "0x1", "relation": "==", "val": "0x1", "register": 4},
           {"name": "x", "qubits": [3], "conditional": 4},
           {"name": "reset", "qubits": [1]},
           circuits.u2(0.3, 0.2, 1), circuits.cx(1, 4),
           {"name": "measure", "qubits": [3], "memory": [1], "register": [1]},
           {"name": "measure", "qubits": [4], "memory": [2], "register": [2]},
           {"name": "measure", "qubits": [1], "memory": [3], "register": [3]}]
    exp = circuits.make_experiment("mid", 5, ins, memory_slots=4)
    _same_as_oracle("qasm_simulator", circuits.make_qobj(exp, shots=400, seed_simulator=77, memory=True))
    _same_as_oracle("statevector_simulator", circuits.make_qobj(exp, shots=1, seed_simulator=78))


@pytest.mark.parametrize("name", ["midcircuit_8", "teleport", "reset_midcircuit", "if_statement",
                                  "transpiled_init_reset_cond_counts"])
def test_shot_batched_execution_on_the_gpu(name):
    """Circuits with mid-circuit measure / reset / conditionals: all shots side by side in one (n + b)-qubit device
    vector (qsv_marginal over qubit + shot qubits, qsv_collapse_batch, qsv_apply_masked) -- memory and counts identical
    to the shot-by-shot path and to the reference's recorded output (qasm_simulator.py:512-614)."""
    import copy

    import golden_io
    import helpers
    from qiskit_sdk_py_b200 import _lib, simulator

    case = golden_io.load_case(name)
    batched = simulator.QasmSimulatorB200()
    before = _lib.launch_count()
    res_b = batched.run(copy.deepcopy(case["qobj"]), **case.get("options", {}))
    launches_b = _lib.launch_count() - before
    assert batched.last_stats["shot_batch_bits"] >= 3
    serial = simulator.QasmSimulatorB200()
    serial.MIN_BATCH_BITS = 99
    before = _lib.launch_count()
    res_s = serial.run(copy.deepcopy(case["qobj"]), **case.get("options", {}))
    launches_s = _lib.launch_count() - before
    assert res_b["results"][0]["data"] == res_s["results"][0]["data"]
    helpers.compare_result(res_b, case)
    assert launches_b * 4 < launches_s, (launches_b, launches_s)  # a few launches per batch instead of per shot


def test_shot_batched_wide_circuit_against_oracle():
    """A 12-qubit mid-circuit experiment, 600 shots in batches of 512 + 88 (partial last batch)."""
    rng = np.random.RandomState(5)
    n = 12
    ins = [circuits.u3(*rng.uniform(0, np.pi, 3), q) for q in range(n)]
    ins += [circuits.cx(q, q + 1) for q in range(n - 1)]
    ins += [{"name": "measure", "qubits": [n - 1], "memory": [0], "register": [0]},
            {"name": "x", "qubits": [2], "conditional": 0},
            {"name": "reset", "qubits": [5]},
            circuits.u2(0.2, 0.7, 5), circuits.cx(5, 0),
            {"name": "bfunc", "mask": "0x1", "relation": "!=", "val": "0x1", "register": 3, "memory": 2},
            {"name": "u1", "qubits": [7], "params": [0.4], "conditional": 3}]
    ins += [{"name": "measure", "qubits": [q], "memory": [3 + q], "register": [4 + q]} for q in range(n)]
    exp = circuits.make_experiment("mid12", n, ins, memory_slots=3 + n)
    _same_as_oracle("qasm_simulator", circuits.make_qobj(exp, shots=600, seed_simulator=5, memory=True))
