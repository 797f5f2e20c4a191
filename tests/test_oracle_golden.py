"""Pin the oracle: the numpy restatement must reproduce the reference's recorded outputs
(tests/golden/*.json, produced by tests/golden/make_golden.py from the unmodified reference),
bit-exact for counts / memory and to 1e-13 for amplitudes."""
import copy
import os

import numpy as np
import pytest

import golden_io
import helpers
from oracle import basicaer_oracle as orc

ALL = golden_io.list_cases()
SMALL = [n for n in ALL if golden_io.load_case(n)["qobj"]["config"]["n_qubits"] <= 14]
LARGE = [n for n in ALL if n not in SMALL]


def oracle_factory(backend):
    if backend == "qasm_simulator":
        return orc.OracleQasmSimulator()
    if backend == "statevector_simulator":
        return orc.OracleQasmSimulator(show_final_state=True)
    return orc.OracleUnitarySimulator()


@pytest.mark.parametrize("name", SMALL)
def test_oracle_matches_reference_golden(name):
    case = golden_io.load_case(name)
    result = helpers.run_case_with(oracle_factory, case)
    helpers.compare_result(result, case, amp_tol=1e-13)


@pytest.mark.slow
@pytest.mark.parametrize("name", [n for n in LARGE if "16" in n or "18" in n])
def test_oracle_matches_reference_golden_medium(name):
    case = golden_io.load_case(name)
    result = helpers.run_case_with(oracle_factory, case)
    helpers.compare_result(result, case, amp_tol=1e-13)


def test_bell_known_answer():
    """BASELINE config 1 / SURVEY 8(c): seed 42, 1024 shots -> {'00': 511, '11': 513}."""
    case = golden_io.load_case("bell_seed42")
    assert case["expected"][0]["counts"] == {"0x0": 511, "0x3": 513}
    res = oracle_factory("qasm_simulator").run(copy.deepcopy(case["qobj"]))
    assert res["results"][0]["data"]["counts"] == {"0x0": 511, "0x3": 513}


def test_index_convention():
    """out[i] = sum_c G[r,c] in[i with bits at qubits[j] <- bit j of c]."""
    rng = np.random.RandomState(0)
    n = 5
    psi = rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)
    for qubits in ([0], [3], [1, 3], [3, 1], [4, 0, 2]):
        k = len(qubits)
        g = rng.standard_normal((2 ** k, 2 ** k)) + 1j * rng.standard_normal((2 ** k, 2 ** k))
        out = orc.apply_unitary(psi.reshape([2] * n), g, qubits).reshape(-1)
        exp = np.zeros_like(psi)
        for i in range(2 ** n):
            r = sum(((i >> q) & 1) << j for j, q in enumerate(qubits))
            for c in range(2 ** k):
                src = i
                for j, q in enumerate(qubits):
                    src = (src & ~(1 << q)) | (((c >> j) & 1) << q)
                exp[i] += g[r, c] * psi[src]
        assert np.max(np.abs(out - exp)) < 1e-13


@pytest.mark.skipif(not os.path.isdir("/root/reference/qiskit"), reason="reference tree not present")
def test_oracle_matches_live_reference():
    """When the reference is present (build container), run it live on a fresh random circuit."""
    import sys
    import warnings

    sys.path.insert(0, os.path.join(helpers.__file__.rsplit("/tests/", 1)[0], "oracle", "ref_env"))
    import ref_import

    ref_import.import_reference()
    warnings.simplefilter("ignore")
    from qiskit import BasicAer
    from qiskit.qobj import QasmQobj
    from qiskit_sdk_py_b200 import circuits as gen

    qobj = gen.random_circuit(7, 9, 321, shots=777, seed_simulator=5, memory=True)
    ref = BasicAer.get_backend("qasm_simulator").run(QasmQobj.from_dict(copy.deepcopy(qobj))).result()
    got = oracle_factory("qasm_simulator").run(copy.deepcopy(qobj))
    assert ref.results[0].data.to_dict()["memory"] == got["results"][0]["data"]["memory"]
    assert dict(ref.results[0].data.to_dict()["counts"]) == got["results"][0]["data"]["counts"]


def test_expval_oracle_matches_compiled_reference():
    """oracle.expval_pauli vs the reference's own Cython code (exp_value.pyx compiled by
    oracle/build_ref.py, values recorded in tests/golden/expval_pauli.json)."""
    fix = golden_io.load_case("expval_pauli")
    for case in fix["cases"]:
        for label, want in zip(case["labels"], case["values"]):
            got = orc.expval_pauli(case["state"], label)
            assert abs(got - want) < 1e-12, (case["n"], label, got, want)
