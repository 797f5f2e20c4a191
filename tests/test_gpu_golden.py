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


@pytest.mark.parametrize("n", [26, 28])
def test_large_properties(n):
    """Sizes beyond the reference's 24-qubit cap: norm preservation and the inverse-circuit echo
    (SURVEY 8(d) config 4 correctness plan)."""
    from qiskit_sdk_py_b200.engine import DeviceState
    from qiskit_sdk_py_b200 import lower

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
