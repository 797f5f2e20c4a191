#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py            # small cases (seconds)
    python tests/golden/make_golden.py --large    # + QFT-20, QV-20, QV-24 (several minutes)

Each fixture is one JSON file holding {"backend", "qobj", "options", "expected"}:
``qobj`` is the Qobj wire dict that was fed to the reference backend
(``backend.run(QasmQobj.from_dict(qobj), **options)``) and ``expected`` the outputs the
reference produced: per experiment ``counts`` (hex keys, insertion-ordered), ``memory``,
``statevector`` / ``unitary`` (bit-exact complex128) and ``seed_simulator``.  For large
states only a strided sample of amplitudes plus the norm and a SHA-256 of the rounded vector
are kept.

Two sources of circuits:
  * circuits restated from the reference's own tests (test/python/basicaer/*.py), built with
    the reference's QuantumCircuit in basis gates and serialised by the reference's
    ``assemble`` -- this pins the wire format (conditionals -> bfunc, register fields, headers);
  * circuits from ``qiskit_sdk_py_b200.circuits`` (QFT / quantum volume / random in basis
    gates) -- the generator output is first checked against the reference's own
    ``QuantumVolume`` / ``QFT`` classes.
"""
import argparse
import copy
import hashlib
import math
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_env"))

import ref_import  # noqa: E402

ref_import.import_reference()
warnings.simplefilter("ignore")

from qiskit import BasicAer, ClassicalRegister, QuantumCircuit, QuantumRegister, assemble  # noqa: E402
from qiskit.qobj import QasmQobj  # noqa: E402

import golden_io  # noqa: E402
from qiskit_sdk_py_b200 import circuits as gen  # noqa: E402

PI = math.pi
SAMPLE_STRIDE = 4099  # prime stride for amplitude samples of large states


def h(circ, q):
    circ.u2(0, PI, q)


def run_reference(backend_name, qobj_dict, options):
    backend = BasicAer.get_backend(backend_name)
    qobj = QasmQobj.from_dict(copy.deepcopy(qobj_dict))
    result = backend.run(qobj, **options).result()
    out = []
    for res in result.results:
        data = res.data.to_dict()
        exp = {"shots": res.shots}
        if hasattr(res, "seed_simulator"):
            exp["seed_simulator"] = res.seed_simulator
        if "counts" in data:
            exp["counts"] = dict(data["counts"])
            exp["counts_order"] = list(data["counts"].keys())
        if "memory" in data:
            exp["memory"] = list(data["memory"])
        if "statevector" in data:
            exp["statevector"] = np.asarray(data["statevector"], dtype=complex)
        if "unitary" in data:
            exp["unitary"] = np.asarray(data["unitary"], dtype=complex)
        out.append(exp)
    return out, result


def compress_large(expected):
    for exp in expected:
        sv = exp.get("statevector")
        if sv is not None and sv.size > (1 << 14):
            idx = np.arange(0, sv.size, SAMPLE_STRIDE, dtype=np.int64)
            exp["statevector_sample_idx"] = idx
            exp["statevector_sample"] = sv[idx].copy()
            exp["statevector_norm2"] = float(np.sum(np.abs(sv) ** 2))
            exp["statevector_sum"] = complex(np.sum(sv))
            exp["statevector_sha256_round9"] = hashlib.sha256(
                np.round(sv.view(np.float64), 9).tobytes()).hexdigest()
            del exp["statevector"]
    return expected


def save(name, backend, qobj_dict, options=None):
    options = options or {}
    expected, _ = run_reference(backend, qobj_dict, options)
    golden_io.dump_case(name, {"backend": backend, "qobj": qobj_dict, "options": options,
                               "expected": compress_large(expected)})
    print("wrote", name)


def assembled(circs, **kw):
    return assemble(circs, **kw).to_dict()


# --------------------------------------------------------------------------------------
def cases_from_reference_tests():
    # README Bell, BASELINE config 1 (README.md:32-41): expect {'00': 511, '11': 513}
    qc = QuantumCircuit(2, 2)
    h(qc, 0)
    qc.cx(0, 1)
    qc.measure([0, 1], [0, 1])
    save("bell_seed42", "qasm_simulator", assembled(qc, shots=1024, seed_simulator=42))
    save("bell_memory", "qasm_simulator", assembled(qc, shots=64, seed_simulator=7, memory=True))
    save("bell_single_shot", "qasm_simulator", assembled(qc, shots=1, seed_simulator=3))
    # statevector: final measures collapse (test_statevector_simulator.py:46-59)
    save("sv_bell_measure_collapse", "statevector_simulator", assembled(qc, seed_simulator=11))
    save("sv_bell_measure_collapse_b", "statevector_simulator", assembled(qc, seed_simulator=12))

    # test_qasm_simulator.py:82-98 repeated qubits
    qr = QuantumRegister(2, "qr")
    cr = ClassicalRegister(4, "cr")
    c = QuantumCircuit(qr, cr)
    c.x(qr[1])
    c.measure(qr[0], cr[0])
    c.measure(qr[1], cr[1])
    c.measure(qr[1], cr[2])
    c.measure(qr[0], cr[3])
    save("sampler_repeated_qubits", "qasm_simulator", assembled(c, shots=100, seed_simulator=73846087))

    # :100-115 single qubit
    qr = QuantumRegister(5, "qr")
    cr = ClassicalRegister(1, "cr")
    circs = []
    for q in range(5):
        c = QuantumCircuit(qr, cr)
        c.x(qr[q])
        c.measure(qr[q], cr[0])
        circs.append(c)
    save("sampler_single_qubit", "qasm_simulator", assembled(circs, shots=100, seed_simulator=73846087))

    # :117-138 partial qubits with barriers
    cr = ClassicalRegister(4, "cr")
    c = QuantumCircuit(qr, cr)
    c.x(qr[3])
    c.x(qr[1])
    c.barrier(qr)
    c.measure(qr[3], cr[1])
    c.barrier(qr)
    c.measure(qr[1], cr[0])
    c.barrier(qr)
    c.measure(qr[0], cr[2])
    c.barrier(qr)
    c.measure(qr[3], cr[3])
    save("sampler_partial_qubit", "qasm_simulator", assembled(c, shots=100, seed_simulator=73846087))

    # partial measurement of a superposition: marginal over a qubit subset
    qr = QuantumRegister(5, "qr")
    cr = ClassicalRegister(3, "cr")
    c = QuantumCircuit(qr, cr)
    for q in range(5):
        c.u3(0.3 + 0.4 * q, 0.1 * q, 0.7, qr[q])
    c.cx(qr[0], qr[3])
    c.cx(qr[4], qr[1])
    c.measure(qr[3], cr[0])
    c.measure(qr[1], cr[2])
    c.measure(qr[4], cr[1])
    save("sampler_marginal_subset", "qasm_simulator",
         assembled(c, shots=4000, seed_simulator=5, memory=True))

    # :158-193 conditionals
    qr = QuantumRegister(3, "qr")
    cr = ClassicalRegister(3, "cr")
    t = QuantumCircuit(qr, cr)
    t.x(qr[0])
    t.x(qr[1])
    t.measure(qr[0], cr[0])
    t.measure(qr[1], cr[1])
    t.x(qr[2]).c_if(cr, 0x3)
    t.measure(qr[0], cr[0])
    t.measure(qr[1], cr[1])
    t.measure(qr[2], cr[2])
    f = QuantumCircuit(qr, cr)
    f.x(qr[0])
    f.measure(qr[0], cr[0])
    f.measure(qr[1], cr[1])
    f.x(qr[2]).c_if(cr, 0x3)
    f.measure(qr[0], cr[0])
    f.measure(qr[1], cr[1])
    f.measure(qr[2], cr[2])
    save("if_statement", "qasm_simulator", assembled([t, f], shots=100, seed_simulator=73846087))

    # :195-238 teleport (h -> u2(0,pi), ry(t) -> u3(t,0,0), z -> u1(pi))
    qr = QuantumRegister(3, "qr")
    cr0 = ClassicalRegister(1, "cr0")
    cr1 = ClassicalRegister(1, "cr1")
    cr2 = ClassicalRegister(1, "cr2")
    c = QuantumCircuit(qr, cr0, cr1, cr2, name="teleport")
    h(c, qr[1])
    c.cx(qr[1], qr[2])
    c.u3(PI / 4, 0, 0, qr[0])
    c.cx(qr[0], qr[1])
    h(c, qr[0])
    c.barrier(qr)
    c.measure(qr[0], cr0[0])
    c.measure(qr[1], cr1[0])
    c.u1(PI, qr[2]).c_if(cr0, 1)
    c.x(qr[2]).c_if(cr1, 1)
    c.measure(qr[2], cr2[0])
    save("teleport", "qasm_simulator", assembled(c, shots=2000, seed_simulator=73846087, memory=True))

    # mid-circuit reset + measure, per-shot RNG order (qasm_simulator.py:255-273, :519-628)
    qr = QuantumRegister(3, "qr")
    cr = ClassicalRegister(3, "cr")
    c = QuantumCircuit(qr, cr)
    h(c, qr[0])
    c.cx(qr[0], qr[1])
    c.u3(1.1, 0.2, 0.3, qr[2])
    c.reset(qr[0])
    h(c, qr[0])
    c.measure(qr[0], cr[0])
    c.reset(qr[2])
    c.u3(0.4, 0.5, 0.6, qr[2])
    c.cx(qr[2], qr[0])
    c.measure(qr[1], cr[1])
    c.measure(qr[2], cr[2])
    save("reset_midcircuit", "qasm_simulator", assembled(c, shots=300, seed_simulator=99, memory=True))
    save("sv_reset_midcircuit", "statevector_simulator", assembled(c, seed_simulator=99))

    # :240-260 memory with two registers
    qr = QuantumRegister(4, "qr")
    cr0 = ClassicalRegister(2, "cr0")
    cr1 = ClassicalRegister(2, "cr1")
    c = QuantumCircuit(qr, cr0, cr1)
    h(c, qr[0])
    c.cx(qr[0], qr[1])
    c.x(qr[3])
    c.measure(qr[0], cr0[0])
    c.measure(qr[1], cr0[1])
    c.measure(qr[2], cr1[0])
    c.measure(qr[3], cr1[1])
    save("memory_two_registers", "qasm_simulator", assembled(c, shots=50, seed_simulator=1, memory=True))

    # :262-285 k-qubit unitary X^(x)k
    x_mat = np.array([[0, 1], [1, 0]])
    circs = []
    for i in range(4):
        nq = i + 1
        multi_x = x_mat
        for _ in range(i):
            multi_x = np.kron(multi_x, x_mat)
        qr = QuantumRegister(nq, "qr")
        cr = ClassicalRegister(nq, "cr")
        c = QuantumCircuit(qr, cr)
        c.unitary(multi_x, qr)
        c.measure(qr, cr)
        circs.append(c)
    save("unitary_multi_x", "qasm_simulator", assembled(circs, shots=1024, seed_simulator=2))

    # random k-qubit unitaries on permuted qubits, statevector (test_statevector_simulator.py:61-83)
    from scipy import stats
    circs = []
    for k, qubits, n in ((1, [2], 4), (2, [3, 0], 4), (3, [4, 0, 2], 5), (3, [0, 1, 2], 3),
                         (4, [1, 5, 2, 4], 6), (5, [6, 0, 3, 1, 5], 7)):
        mat = stats.unitary_group.rvs(2 ** k, random_state=np.random.default_rng(100 + k + n))
        c = QuantumCircuit(n)
        for q in range(n):
            c.u3(0.2 + 0.3 * q, 0.5 * q, 0.1, q)
        c.unitary(mat, qubits)
        circs.append(c)
    save("sv_random_unitaries", "statevector_simulator", assembled(circs))

    # test_multi_registers_convention.py:24-59
    qreg0 = QuantumRegister(2, "q0")
    creg0 = ClassicalRegister(2, "c0")
    qreg1 = QuantumRegister(2, "q1")
    creg1 = ClassicalRegister(2, "c1")
    circ = QuantumCircuit(qreg0, qreg1, creg0, creg1)
    circ.x(qreg0[1])
    circ.x(qreg1[0])
    meas = QuantumCircuit(qreg0, qreg1, creg0, creg1)
    meas.measure(qreg0, creg0)
    meas.measure(qreg1, creg1)
    qc = circ.compose(meas)
    save("multi_registers_counts", "qasm_simulator", assembled(qc, shots=1024, seed_simulator=4))
    save("multi_registers_state", "statevector_simulator", assembled(circ))
    save("multi_registers_unitary", "unitary_simulator", assembled(circ))

    # test_statevector_simulator.py:85-114 global phase
    c = QuantumCircuit(4)
    c.x(range(4))
    c.global_phase = 0.5
    save("sv_global_phase", "statevector_simulator", assembled(c))
    save("unitary_global_phase_x4", "unitary_simulator", assembled(c))

    # test_unitary_simulator.py:35-92 (h -> u2(0,pi); y -> u3(pi,pi/2,pi/2); z -> u1(pi))
    qr = QuantumRegister(3)
    cr = ClassicalRegister(3)
    qc1, qc2, qc3, qc4, qc5 = (QuantumCircuit(qr, cr) for _ in range(5))
    for q in range(3):
        h(qc1, qr[q])
    qc2.cx(qr[0], qr[1])
    qc3.u3(PI, PI / 2, PI / 2, qr[0])
    qc3.cx(qr[1], qr[2])
    h(qc4, qr[0])
    qc4.cx(qr[0], qr[1])
    qc4.cx(qr[1], qr[2])
    qc5.x(qr[0])
    qc5.u3(PI, PI / 2, PI / 2, qr[0])
    qc5.u3(PI, PI / 2, PI / 2, qr[1])
    qc5.u1(PI, qr[1])
    qc5.u1(PI, qr[2])
    qc5.x(qr[2])
    save("unitary_five_circuits", "unitary_simulator", assembled([qc1, qc2, qc3, qc4, qc5]))

    # :117-134 global phase XZH (h; z; x in basis gates)
    c = QuantumCircuit(1)
    h(c, 0)
    c.u1(PI, 0)
    c.x(0)
    save("unitary_xzh", "unitary_simulator", assembled(c))

    # random unitaries on the unitary simulator (:94-115)
    circs = []
    for k, qubits, n in ((1, [1], 3), (2, [2, 0], 3), (3, [1, 3, 0], 4)):
        mat = stats.unitary_group.rvs(2 ** k, random_state=np.random.default_rng(7 + k))
        c = QuantumCircuit(n)
        c.u3(0.3, 0.2, 0.1, 0)
        c.cx(0, n - 1)
        c.unitary(mat, qubits)
        c.rz(0.7, 1)
        c.sx(0)
        circs.append(c)
    save("unitary_random", "unitary_simulator", assembled(circs))

    # initial_statevector option (qasm_simulator.py:299-311) and chop_threshold
    rng = np.random.RandomState(5)
    init = rng.standard_normal(8) + 1j * rng.standard_normal(8)
    init /= np.linalg.norm(init)
    c = QuantumCircuit(3)
    h(c, 1)
    c.cx(1, 2)
    c.rz(0.3, 0)
    save("sv_initial_statevector", "statevector_simulator", assembled(c),
         {"initial_statevector": init})
    c2 = QuantumCircuit(3, 3)
    h(c2, 1)
    c2.cx(1, 2)
    c2.measure(range(3), range(3))
    save("qasm_initial_statevector", "qasm_simulator",
         assembled(c2, shots=500, seed_simulator=8), {"initial_statevector": init})
    # initial_unitary
    iu = stats.unitary_group.rvs(4, random_state=np.random.default_rng(3))
    c = QuantumCircuit(2)
    h(c, 0)
    c.cx(0, 1)
    save("unitary_initial_unitary", "unitary_simulator", assembled(c), {"initial_unitary": iu})


def check_generators_against_reference():
    """The qiskit-free generators must emit what the reference's classes emit."""
    from qiskit.circuit.library import QFT, QuantumVolume
    from qiskit.quantum_info import Statevector

    for n in (4, 7):
        qv = QuantumVolume(n, n, seed=1234)
        inner = qv.data[0][0].definition
        ours = gen.quantum_volume_instructions(n, n, 1234)
        assert len(inner.data) == len(ours), (len(inner.data), len(ours))
        for (inst, qargs, _), op in zip(inner.data, ours):
            assert [inner.qubits.index(q) for q in qargs] == op["qubits"]
            assert np.array_equal(inst.to_matrix(), op["params"][0])
    print("quantum_volume generator == reference QuantumVolume (n=4,7)")
    # QFT: compare the statevector of our basis-gate stream with quantum_info's evolution of QFT(n)
    for n in (3, 6):
        qobj = gen.qft(n, state_seed=9)
        exp, _ = run_reference("statevector_simulator", qobj, {})
        qc = QuantumCircuit(n)
        for ins in gen.product_state_layer(n, 9):
            qc.u3(*ins["params"], ins["qubits"][0])
        qc.compose(QFT(n), inplace=True)
        ref = Statevector.from_instruction(qc).data
        err = np.max(np.abs(ref - exp[0]["statevector"]))
        assert err < 1e-12, err
    print("qft generator == Statevector(QFT(n)) (n=3,6)")


def cases_from_generators():
    save("qft_5", "statevector_simulator", gen.qft(5))
    save("qft_10_product", "statevector_simulator", gen.qft(10, state_seed=3))
    save("qft_12_counts", "qasm_simulator",
         gen.qft(12, state_seed=4, shots=4096, measure_final=True, seed_simulator=21))
    for n in (2, 3, 5, 8, 11, 13):
        save("qv_%d_counts" % n, "qasm_simulator", gen.quantum_volume(n, shots=2048, seed_simulator=42))
        save("qv_%d_state" % n, "statevector_simulator",
             gen.quantum_volume(n, measure_final=False, shots=1))
    for n, depth, seed in ((1, 6, 1), (2, 6, 2), (3, 8, 3), (4, 8, 4), (6, 10, 5), (9, 12, 6),
                           (12, 12, 7), (14, 10, 8)):
        save("random_%d_state" % n, "statevector_simulator",
             gen.random_circuit(n, depth, seed, measure_final=False, shots=1))
        save("random_%d_counts" % n, "qasm_simulator",
             gen.random_circuit(n, depth, seed, shots=1500, seed_simulator=seed + 100, memory=True))
    # unitary simulator on generated streams
    for n in (2, 4, 6):
        q = gen.random_circuit(n, 6, 40 + n, measure_final=False, shots=1)
        save("random_%d_unitary" % n, "unitary_simulator", q)


def expval_cases():
    """Pauli expectation values from the reference's compiled Cython code (oracle/build_ref.py)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import build_ref

    if build_ref.build() is None:
        print("exp_value.pyx not found; skipping expval goldens")
        return
    from qiskit.quantum_info import Pauli, Statevector
    import qiskit.quantum_info.states.statevector as svmod

    assert "oracle/_ref" in svmod.expval_pauli_no_x.__module__ or hasattr(svmod.expval_pauli_no_x, "__call__")
    rng = np.random.RandomState(2024)
    cases = []
    for n in (1, 2, 5, 9, 12):
        vec = rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)
        vec /= np.linalg.norm(vec)
        labels = ["I" * n, "Z" * n, "X" * n, "Y" * n]
        for _ in range(12):
            labels.append("".join(rng.choice(list("IXYZ"), n)))
        sv = Statevector(vec)
        vals = [complex(sv.expectation_value(Pauli(lab))) for lab in labels]
        cases.append({"n": n, "state": vec, "labels": labels, "values": np.array(vals)})
    golden_io.dump_case("expval_pauli", {"backend": "expval", "cases": cases,
                                         "qobj": {"config": {"n_qubits": 12}}})
    print("wrote expval_pauli")


def transpiled_cases():
    """Circuits that go through the reference's own transpile() + assemble() (level 0/1 are deterministic;
    the DAG is backed by oracle/ref_env/retworkx_shim.py), i.e. the Qobj streams execute() really produces."""
    from qiskit import transpile
    from qiskit.circuit.library import QFT, GroverOperator, QuantumVolume

    sim = BasicAer.get_backend("qasm_simulator")

    def lib_gates(n):
        qc = QuantumCircuit(n, n)
        for q in range(n):
            qc.h(q)
        qc.ccx(0, 1, 2)
        qc.swap(1, 3)
        qc.cswap(0, 2, 4)
        qc.crz(0.7, 3, 4)
        qc.ch(2, 0)
        qc.cp(0.4, 1, 2)
        qc.cy(4, 1)
        qc.cz(0, 3)
        qc.rxx(0.3, 1, 4)
        qc.rzz(0.9, 2, 3)
        qc.mcx([0, 1, 2], 3)
        qc.t(0)
        qc.sdg(1)
        qc.sx(2)
        qc.ry(0.2, 3)
        qc.global_phase = 0.5
        return qc

    for level in (0, 1):
        qc = lib_gates(5)
        save("transpiled_lib_gates_l%d_state" % level, "statevector_simulator",
             assembled(transpile(qc, sim, optimization_level=level), shots=1))
        qc.measure(range(5), range(5))
        save("transpiled_lib_gates_l%d_counts" % level, "qasm_simulator",
             assembled(transpile(qc, sim, optimization_level=level), shots=4000, seed_simulator=7))
    # library QFT (with the final swaps) on a product state
    qc = QuantumCircuit(9)
    for q in range(9):
        qc.ry(0.3 + 0.2 * q, q)
    qc = qc.compose(QFT(9))
    save("transpiled_qft_lib_9_state", "statevector_simulator", assembled(transpile(qc, sim, optimization_level=1), shots=1))
    # library quantum volume: SU(4) blocks stay ``unitary`` instructions
    qv = QuantumVolume(10, 10, seed=17)
    qv.measure_all()
    save("transpiled_qv_lib_10_counts", "qasm_simulator",
         assembled(transpile(qv, sim, optimization_level=1), shots=4096, seed_simulator=42))
    # Grover iteration (multi-controlled gates unrolled to the basis)
    oracle_c = QuantumCircuit(5)
    oracle_c.x([1, 3])
    oracle_c.h(4)
    oracle_c.mcx([0, 1, 2, 3], 4)
    oracle_c.h(4)
    oracle_c.x([1, 3])
    grover = QuantumCircuit(5, 5)
    grover.h(range(5))
    op = GroverOperator(oracle_c)
    for _ in range(3):
        grover = grover.compose(op)
    grover.measure(range(5), range(5))
    save("transpiled_grover_5_counts", "qasm_simulator",
         assembled(transpile(grover, sim, optimization_level=1), shots=2000, seed_simulator=5))
    # initialize + reset + conditional through the transpiler
    rng = np.random.RandomState(3)
    vec = rng.randn(8) + 1j * rng.randn(8)
    vec /= np.linalg.norm(vec)
    qr = QuantumRegister(4, "qr")
    cr = ClassicalRegister(4, "cr")
    qc = QuantumCircuit(qr, cr)
    qc.h(3)
    qc.initialize(vec, [0, 1, 2])
    qc.cx(3, 0)
    qc.measure(0, 0)
    qc.reset(1)
    qc.h(1)
    qc.x(2).c_if(cr, 1)
    qc.measure(range(4), range(4))
    save("transpiled_init_reset_cond_counts", "qasm_simulator",
         assembled(transpile(qc, sim, optimization_level=1), shots=3000, seed_simulator=21, memory=True))


def midcircuit_wide_cases():
    """Mid-circuit measure / reset / conditionals on registers wide enough for a state sharded over 8 ranks (the
    reference's own mid-circuit tests have 3-4 qubits): per-shot simulation with the exact RNG order
    (qasm_simulator.py:512-531, :561-614).  Used by the multi-GPU parity gate of bench.py and tests/test_gpu_dist.py."""
    rng = np.random.RandomState(2024)

    def build(n, name):
        qr = QuantumRegister(n, "q")
        ca = ClassicalRegister(3, "ca")
        cb = ClassicalRegister(n, "cb")
        c = QuantumCircuit(qr, ca, cb, name=name)
        for q in range(n):
            c.u3(*rng.uniform(0, PI, 3), qr[q])
        for q in range(n - 1):
            c.cx(qr[q], qr[q + 1])
        for q in range(n):
            c.u3(*rng.uniform(0, PI, 3), qr[q])
        c.cx(qr[n - 1], qr[0])                     # highest (global on every sharding) -> lowest
        c.measure(qr[n - 1], ca[0])                # mid-circuit measure of a global qubit
        c.u1(0.7, qr[n - 2]).c_if(ca, 1)
        c.x(qr[1]).c_if(ca, 1)
        c.reset(qr[n - 2])                         # reset of a global qubit
        h(c, qr[n - 2])
        c.cx(qr[n - 2], qr[2])
        c.measure(qr[2], ca[1])
        c.u3(0.3, 0.2, 0.1, qr[n - 1]).c_if(ca, 3)
        c.reset(qr[0])
        c.u2(0.1, 0.4, qr[0])
        c.cx(qr[0], qr[n - 1])
        c.measure(qr[n - 3], ca[2])
        for q in range(n):
            c.measure(qr[q], cb[q])
        return c

    save("midcircuit_8", "qasm_simulator", assembled(build(8, "midcircuit_8"), shots=150, seed_simulator=31, memory=True))
    save("sv_midcircuit_10", "statevector_simulator", assembled(build(10, "sv_midcircuit_10"), seed_simulator=77))


def large_cases():
    save("qft_16_product", "statevector_simulator", gen.qft(16, state_seed=3))
    save("qv_16_counts", "qasm_simulator", gen.quantum_volume(16, shots=8192, seed_simulator=42))
    save("qv_18_state", "statevector_simulator", gen.quantum_volume(18, measure_final=False, shots=1))
    save("qft_20_product", "statevector_simulator", gen.qft(20, state_seed=3))  # BASELINE config 2
    save("qft_20", "statevector_simulator", gen.qft(20))
    save("qv_20_counts", "qasm_simulator", gen.quantum_volume(20, shots=8192, seed_simulator=42))


def huge_cases():
    # BASELINE config 3: ~4 minutes on one core
    save("qv_24_counts", "qasm_simulator", gen.quantum_volume(24, shots=8192, seed_simulator=42))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--large", action="store_true")
    ap.add_argument("--huge", action="store_true")
    ap.add_argument("--only-large", action="store_true")
    ap.add_argument("--only-expval", action="store_true")
    ap.add_argument("--only-transpiled", action="store_true")
    ap.add_argument("--only-midcircuit-wide", action="store_true")
    args = ap.parse_args()
    if args.only_midcircuit_wide:
        midcircuit_wide_cases()
        sys.exit(0)
    if args.only_expval:
        expval_cases()
        sys.exit(0)
    if args.only_transpiled:
        transpiled_cases()
        sys.exit(0)
    if not args.only_large:
        check_generators_against_reference()
        cases_from_reference_tests()
        cases_from_generators()
        expval_cases()
        transpiled_cases()
        midcircuit_wide_cases()
    if args.large or args.only_large:
        large_cases()
    if args.huge:
        huge_cases()
