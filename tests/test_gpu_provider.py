"""The reference's plugin path ON THE GPU: execute(circuit, B200Aer backend) -> BackendV1.run -> assemble ->
QasmQobj.to_dict -> simulator.py -> libqsv (CUDA) -> Result.from_dict, compared with BasicAer IN THE SAME PROCESS
(execute_function.py:335-377, qasm_simulator.py:376-424, qiskit/test/providers/backend.py:53-75).

The reference package is imported from /root/reference (build container) or from the copy that
__graft_entry__.build() leaves under baseline/_ref/ (it ships to the GPU box with the working tree), through the
test-only import shims of oracle/ref_env.  Skipped only if neither exists."""
import os
import sys
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_env"))
import ref_import  # noqa: E402


@pytest.fixture(scope="module")
def qk():
    if not ref_import.reference_available():
        pytest.skip("no reference package (neither /root/reference nor baseline/_ref)")
    ref_import.import_reference()
    warnings.simplefilter("ignore")
    import qiskit

    return qiskit


@pytest.fixture(scope="module")
def b200(qk):
    from qiskit_sdk_py_b200 import _lib
    from qiskit_sdk_py_b200.provider import B200Aer

    _lib.load()  # the CUDA library must be the thing that runs: no fallback exists
    return B200Aer


def test_readme_bell_through_execute(qk, b200):
    """README.md:32-41 literally: h, cx, measure; execute() transpiles to the backend's basis."""
    qc = qk.QuantumCircuit(2, 2)
    qc.h(0)
    qc.cx(0, 1)
    qc.measure([0, 1], [0, 1])
    before = _launches()
    ours = qk.execute(qc, b200.get_backend("qasm_simulator"), shots=1024, seed_simulator=42).result()
    assert _launches() > before, "execute() on the B200 backend must launch libqsv kernels"
    ref = qk.execute(qc, qk.BasicAer.get_backend("qasm_simulator"), shots=1024, seed_simulator=42).result()
    assert ours.get_counts() == ref.get_counts() == {"00": 511, "11": 513}
    assert ours.results[0].header.to_dict() == ref.results[0].header.to_dict()
    assert ours.backend_name == "qasm_simulator" and ours.success


def _launches():
    from qiskit_sdk_py_b200 import _lib

    return _lib.launch_count()


def test_transpiled_qft_statevector(qk, b200):
    from qiskit.circuit.library import QFT

    qc = qk.QuantumCircuit(10)
    for q in range(10):
        qc.ry(0.1 + 0.2 * q, q)
    qc.append(QFT(10), range(10))
    for level in (0, 1, 3):
        ours = qk.execute(qc, b200.get_backend("statevector_simulator"), optimization_level=level).result()
        ref = qk.execute(qc, qk.BasicAer.get_backend("statevector_simulator"), optimization_level=level).result()
        a, b = np.asarray(ours.get_statevector()), np.asarray(ref.get_statevector())
        assert a.dtype == np.complex128 and np.max(np.abs(a - b)) < 1e-10


def test_conditional_and_reset_circuit(qk, b200):
    qr = qk.QuantumRegister(4, "qr")
    ca = qk.ClassicalRegister(2, "ca")
    cb = qk.ClassicalRegister(4, "cb")
    qc = qk.QuantumCircuit(qr, ca, cb)
    qc.h(qr[0])
    qc.cx(qr[0], qr[3])
    qc.ry(1.1, qr[1])
    qc.measure(qr[3], ca[0])
    qc.x(qr[2]).c_if(ca, 1)
    qc.reset(qr[0])
    qc.h(qr[0])
    qc.measure(qr[1], ca[1])
    qc.z(qr[0]).c_if(ca, 3)
    qc.measure(qr, cb)
    kw = dict(shots=400, seed_simulator=17, memory=True)
    ours = qk.execute(qc, b200.get_backend("qasm_simulator"), **kw).result()
    ref = qk.execute(qc, qk.BasicAer.get_backend("qasm_simulator"), **kw).result()
    assert ours.get_memory() == ref.get_memory()
    assert ours.get_counts() == ref.get_counts()


def test_backend_run_variants_and_unitary(qk, b200):
    """run(circuit), run([circuits]), run(qobj) + the unitary backend; options and errors as the reference's."""
    from qiskit.providers.basicaer import BasicAerError

    qc = qk.QuantumCircuit(3, 3)
    qc.u3(0.3, 0.2, 0.1, 0)
    qc.cx(0, 2)
    qc.sx(1)
    qc.measure([0, 1, 2], [0, 1, 2])
    ours_b, ref_b = b200.get_backend("qasm_simulator"), qk.BasicAer.get_backend("qasm_simulator")
    res = ours_b.run([qc, qc], shots=300, seed_simulator=5, memory=True).result()
    ref = ref_b.run([qc, qc], shots=300, seed_simulator=5, memory=True).result()
    assert res.get_memory(1) == ref.get_memory(1) and res.get_counts(0) == ref.get_counts(0)
    qobj = qk.assemble(qc, shots=200, seed_simulator=9)
    with pytest.warns(PendingDeprecationWarning):
        assert ours_b.run(qobj).result().get_counts() == ref_b.run(qobj).result().get_counts()
    qu = qk.QuantumCircuit(4)
    qu.h(3)
    qu.cx(3, 0)
    qu.t(1)
    qu.global_phase = 0.3
    un = qk.execute(qu, b200.get_backend("unitary_simulator")).result().get_unitary()
    un_ref = qk.execute(qu, qk.BasicAer.get_backend("unitary_simulator")).result().get_unitary()
    assert np.max(np.abs(np.asarray(un) - np.asarray(un_ref))) < 1e-10
    with pytest.raises(BasicAerError, match="Unsupported"):
        b200.get_backend("unitary_simulator").run(qc)
    sv_b = b200.get_backend("statevector_simulator")
    sv_b.set_options(initial_statevector=np.array([0, 1, 0, 0], dtype=complex))
    try:
        q2 = qk.QuantumCircuit(2)
        q2.rz(0.4, 0)
        got = sv_b.run(q2).result().get_statevector()
        assert abs(abs(got[1]) - 1.0) < 1e-12
    finally:
        sv_b.set_options(initial_statevector=None)


def test_k_qubit_unitary_instruction(qk, b200):
    """test_qasm_simulator.py:262-285: a k-qubit `unitary` instruction (here k = 3 and 5) straight to the backend."""
    from qiskit.quantum_info import random_unitary

    for k, n in ((3, 5), (5, 8)):
        qc = qk.QuantumCircuit(n, n)
        for q in range(n):
            qc.u3(0.2 * (q + 1), 0.1, 0.3, q)
        qc.unitary(random_unitary(2 ** k, seed=k), list(range(n - 1, n - 1 - k, -1)))
        qc.measure(range(n), range(n))
        kw = dict(shots=1000, seed_simulator=3)
        ours = b200.get_backend("qasm_simulator").run(qc, **kw).result()
        ref = qk.BasicAer.get_backend("qasm_simulator").run(qc, **kw).result()
        assert ours.get_counts() == ref.get_counts()
