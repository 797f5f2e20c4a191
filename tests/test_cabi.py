"""The C-ABI library loads and exports every symbol include/qsv.h declares (no compute calls: there is
no GPU in the CPU test environment)."""
import ctypes
import os
import re

import pytest

from qiskit_sdk_py_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "qsv.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qsv_[a-z0-9_]+)\s*\(", text)))


def test_library_built():
    assert os.path.exists(_lib.LIB_PATH), "run `python __graft_entry__.py` to build libqsv.so"


def test_every_declared_symbol_is_exported_and_bound():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 20
    for name in syms:
        assert hasattr(lib, name), "libqsv.so does not export %s" % name
        assert name in _lib.SIGNATURES, "no ctypes signature for %s" % name
    assert set(_lib.SIGNATURES) == set(syms)


def test_version_and_error_string():
    lib = _lib.load()
    assert lib.qsv_version() == 100
    assert isinstance(lib.qsv_last_error(), bytes)
    assert lib.qsv_reduce_workspace_bytes() > 0
    assert lib.qsv_sample_workspace_bytes(20, 1024) > 8 * 1024
    assert lib.qsv_marginal_workspace_bytes(20, 3) >= 8 * 8


def test_argument_validation_without_gpu():
    """Argument errors are reported before any CUDA call."""
    lib = _lib.load()
    rc = lib.qsv_apply_pass(None, 4, None, 0, None, 0, None)
    assert rc == _lib.QSV_EINVAL
    assert b"qsv_apply_pass" in lib.qsv_last_error()
    assert lib.qsv_prob_qubit(None, 4, 0, None, None, None) == _lib.QSV_EINVAL


def test_no_cpu_fallback_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from qiskit_sdk_py_b200 import circuits, simulator

    with pytest.raises(_lib.QsvLibraryError, match="no CPU fallback"):
        simulator.QasmSimulatorB200(max_qubits=24).run(circuits.bell())


def test_engine_does_not_import_oracle():
    """Product code must not reach into oracle/ (test infrastructure)."""
    pkg = os.path.join(ROOT, "qiskit_sdk_py_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn
