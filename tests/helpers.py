"""Shared comparison helpers for the parity tests."""
import copy

import numpy as np

import golden_io

AMP_TOL = 1e-10  # north_star: max-abs amplitude error in complex128


def run_case_with(sim_factory, case):
    """Run a golden case's qobj through ``sim_factory(backend_name)`` -> result dict."""
    sim = sim_factory(case["backend"])
    return sim.run(copy.deepcopy(case["qobj"]), **case.get("options", {}))


def compare_result(result, case, amp_tol=AMP_TOL, exact_counts=True):
    """Compare a result dict with the reference's recorded outputs."""
    assert len(result["results"]) == len(case["expected"])
    for got, exp in zip(result["results"], case["expected"]):
        data = got["data"]
        assert got["shots"] == exp["shots"]
        if "seed_simulator" in exp and "seed_simulator" in case["qobj"]["config"]:
            assert got["seed_simulator"] == exp["seed_simulator"]
        if "counts" in exp:
            if exact_counts:
                assert dict(data["counts"]) == exp["counts"]
                assert list(data["counts"].keys()) == exp["counts_order"]
            else:
                assert sum(data["counts"].values()) == sum(exp["counts"].values())
        else:
            assert "counts" not in data
        if "memory" in exp:
            assert list(data["memory"]) == exp["memory"]
        else:
            assert "memory" not in data
        if "statevector" in exp:
            sv = np.asarray(data["statevector"])
            assert sv.dtype == np.complex128 and sv.shape == exp["statevector"].shape
            assert np.max(np.abs(sv - exp["statevector"])) <= amp_tol
            # chopped entries are exact zeros in both
            assert np.array_equal(sv == 0, exp["statevector"] == 0) or amp_tol > 1e-13
        if "statevector_sample" in exp:
            sv = np.asarray(data["statevector"])
            idx = exp["statevector_sample_idx"]
            assert np.max(np.abs(sv[idx] - exp["statevector_sample"])) <= amp_tol
            assert abs(float(np.sum(np.abs(sv) ** 2)) - exp["statevector_norm2"]) <= 1e-9
            assert abs(complex(np.sum(sv)) - exp["statevector_sum"]) <= 1e-7
        if "unitary" in exp:
            un = np.asarray(data["unitary"])
            assert un.dtype == np.complex128 and un.shape == exp["unitary"].shape
            assert np.max(np.abs(un - exp["unitary"])) <= amp_tol


def small_cases(max_qubits=14):
    out = []
    for name in golden_io.list_cases():
        case = golden_io.load_case(name)
        if case["qobj"]["config"]["n_qubits"] <= max_qubits:
            out.append(name)
    return out
