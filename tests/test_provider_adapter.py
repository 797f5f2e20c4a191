"""The qiskit plugin adapter (provider.py) against the live reference, on CPU: the device is replaced
by tests/fake_state.FakeState, so this exercises BackendV1/JobV1/ProviderV1 wiring, assemble(),
Qobj -> dict conversion, Result.from_dict and get_counts/get_statevector/get_unitary formatting.
Skipped where the reference tree (and therefore qiskit) is not available, e.g. on the GPU box."""
import os
import sys
import warnings

import numpy as np
import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "qiskit")), reason="reference/qiskit not present")


@pytest.fixture(scope="module")
def qk():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "ref_env"))
    import ref_import

    ref_import.import_reference()
    warnings.simplefilter("ignore")
    import qiskit

    return qiskit


@pytest.fixture(scope="module")
def provider(qk):
    from fake_state import FakeState
    from qiskit_sdk_py_b200.provider import B200Provider

    return B200Provider(simulator_kwargs={"state_factory": FakeState, "max_qubits": 24})


def bell(qk):
    qc = qk.QuantumCircuit(2, 2)
    qc.u2(0, np.pi, 0)
    qc.cx(0, 1)
    qc.measure([0, 1], [0, 1])
    return qc


def test_provider_lists_same_backends(qk, provider):
    names = [b.name() for b in provider.backends()]
    assert names == [b.name() for b in qk.BasicAer.backends()]
    assert provider.get_backend("local_qasm_simulator_py").name() == "qasm_simulator"
    from qiskit.providers.exceptions import QiskitBackendNotFoundError

    with pytest.raises(QiskitBackendNotFoundError):
        provider.get_backend("nope")


def test_backend_conformance(qk, provider):
    """qiskit/test/providers/backend.py:53-75 (BackendTestCase) restated."""
    for ref_backend in qk.BasicAer.backends():
        backend = provider.get_backend(ref_backend.name())
        cfg, ref_cfg = backend.configuration(), ref_backend.configuration()
        assert cfg.backend_name == ref_cfg.backend_name
        assert cfg.basis_gates == ref_cfg.basis_gates
        assert cfg.coupling_map is None and cfg.simulator and cfg.local
        assert cfg.max_shots == ref_cfg.max_shots
        assert backend.properties() is None
        assert backend.status().backend_name == backend.name()
        assert set(vars(backend.options)) == set(vars(ref_backend.options))


def test_run_circuit_counts_match_reference(qk, provider):
    qc = bell(qk)
    backend = provider.get_backend("qasm_simulator")
    ref = qk.BasicAer.get_backend("qasm_simulator")
    job = backend.run(qc, shots=1024, seed_simulator=42)
    assert job.status().name == "DONE"
    res = job.result()
    ref_res = ref.run(qc, shots=1024, seed_simulator=42).result()
    assert res.get_counts() == ref_res.get_counts() == {"00": 511, "11": 513}
    assert res.backend_name == "qasm_simulator" and res.success
    assert res.results[0].header.to_dict() == ref_res.results[0].header.to_dict()
    # list of circuits, memory
    res = backend.run([qc, qc], shots=50, seed_simulator=3, memory=True).result()
    ref_res = ref.run([qc, qc], shots=50, seed_simulator=3, memory=True).result()
    assert res.get_memory(1) == ref_res.get_memory(1)


def test_run_qobj_warns_and_matches(qk, provider):
    qobj = qk.assemble(bell(qk), shots=200, seed_simulator=9)
    backend = provider.get_backend("qasm_simulator")
    with pytest.warns(PendingDeprecationWarning):
        res = backend.run(qobj).result()
    ref_res = qk.BasicAer.get_backend("qasm_simulator").run(qobj).result()
    assert res.get_counts() == ref_res.get_counts()
    assert res.qobj_id == qobj.qobj_id


def test_unknown_option_warns(qk, provider):
    with pytest.warns(UserWarning, match="Option foo is not used by this backend"):
        provider.get_backend("qasm_simulator").run(bell(qk), foo=1)


def test_statevector_and_unitary(qk, provider):
    qc = qk.QuantumCircuit(3)
    qc.u3(0.3, 0.2, 0.1, 0)
    qc.cx(0, 2)
    qc.rz(0.5, 1)
    qc.sx(1)
    qc.global_phase = 0.25
    sv = provider.get_backend("statevector_simulator").run(qc).result().get_statevector()
    ref_sv = qk.BasicAer.get_backend("statevector_simulator").run(qc).result().get_statevector()
    assert np.max(np.abs(sv - ref_sv)) < 1e-13
    un = provider.get_backend("unitary_simulator").run(qc).result().get_unitary()
    ref_un = qk.BasicAer.get_backend("unitary_simulator").run(qc).result().get_unitary()
    assert np.max(np.abs(un - ref_un)) < 1e-13


def test_errors_are_basicaer_errors(qk, provider):
    from qiskit.providers.basicaer import BasicAerError

    qc = qk.QuantumCircuit(30)
    with pytest.raises(BasicAerError, match="Number of qubits 30 is greater than maximum"):
        provider.get_backend("qasm_simulator").run(qc)
    with pytest.raises(BasicAerError, match="Unsupported"):
        provider.get_backend("unitary_simulator").run(bell(qk))


def test_conditional_circuit(qk, provider):
    qr = qk.QuantumRegister(3, "qr")
    cr = qk.ClassicalRegister(3, "cr")
    t = qk.QuantumCircuit(qr, cr)
    t.x(qr[0])
    t.x(qr[1])
    t.measure(qr[0], cr[0])
    t.measure(qr[1], cr[1])
    t.x(qr[2]).c_if(cr, 0x3)
    t.measure(qr[2], cr[2])
    res = provider.get_backend("qasm_simulator").run(t, shots=100, seed_simulator=1).result()
    assert res.get_counts() == {"111": 100}


def test_set_options_persist_like_the_reference(qk, provider):
    """options set with backend.set_options() are the defaults of every later run (the reference reads
    self.options in _set_options, qasm_simulator.py:292-293 / unitary_simulator.py:160-161)."""
    from qiskit_sdk_py_b200.provider import B200Provider
    from fake_state import FakeState

    ours = B200Provider(simulator_kwargs={"state_factory": FakeState, "max_qubits": 24})
    qc = qk.QuantumCircuit(1)
    qc.rz(0.3, 0)
    for name, key, value in (("statevector_simulator", "initial_statevector", np.array([0, 1], dtype=complex)),
                             ("qasm_simulator", "initial_statevector", np.array([0, 1], dtype=complex))):
        b, r = ours.get_backend(name), type(qk.BasicAer.get_backend(name))()
        b.set_options(**{key: value})
        r.set_options(**{key: value})
        if name == "statevector_simulator":
            got = b.run(qc).result().get_statevector()
            want = r.run(qc).result().get_statevector()
            assert np.max(np.abs(got - want)) < 1e-13 and abs(got[1]) > 0.9
            # a per-run option still wins, and the stored one is back on the next run
            got = b.run(qc, initial_statevector=[1, 0]).result().get_statevector()
            assert abs(got[0]) > 0.9
            assert abs(b.run(qc).result().get_statevector()[1]) > 0.9
        else:
            qm = qk.QuantumCircuit(1, 1)
            qm.measure(0, 0)
            assert b.run(qm, shots=20, seed_simulator=1).result().get_counts() == \
                r.run(qm, shots=20, seed_simulator=1).result().get_counts() == {"1": 20}
    b, r = ours.get_backend("statevector_simulator"), type(qk.BasicAer.get_backend("statevector_simulator"))()
    b.set_options(chop_threshold=0.5, initial_statevector=None)  # the same backend object: the option had persisted
    r.set_options(chop_threshold=0.5)
    qh = qk.QuantumCircuit(2)
    qh.u3(0.4, 0, 0, 0)
    assert np.array_equal(b.run(qh).result().get_statevector() == 0, r.run(qh).result().get_statevector() == 0)
    b, r = ours.get_backend("unitary_simulator"), type(qk.BasicAer.get_backend("unitary_simulator"))()
    x = np.array([[0, 1], [1, 0]], dtype=complex)
    b.set_options(initial_unitary=x)
    r.set_options(initial_unitary=x)
    qi = qk.QuantumCircuit(1)
    qi.rz(0.2, 0)
    assert np.max(np.abs(b.run(qi).result().get_unitary() - r.run(qi).result().get_unitary())) < 1e-13
