"""Import the UNMODIFIED reference (Qiskit Terra 0.18, /root/reference) in this container.

TEST INFRASTRUCTURE ONLY.  Used by tests/golden/make_golden.py (to generate the committed
golden fixtures), by the adapter tests (CPU: numpy test double; GPU box: the real CUDA backends next to
BasicAer in one process, from the baseline/_ref copy) and by bench.py's reference arm / cpu_baseline leg, which
time the reference's own QasmSimulatorPy.  Nothing in the product package imports this file.

The reference needs a few third-party modules that are not installed here and cannot be
installed (no network): retworkx (-> retworkx_shim.py), ply (-> qasm2_reader.py), fastjsonschema, python-constraint, plus three Cython
extensions that are not built in the read-only tree.  None of them is on the BasicAer
arithmetic path (qiskit/providers/basicaer/qasm_simulator.py:145-161 is numpy only), so
they are replaced by inert stand-ins registered in ``sys.modules`` *before* ``import qiskit``.
The reference sources are read where they lie; nothing is copied or written.
"""
import importlib
import os
import sys
import types

_REPO_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _default_root():
    """/root/reference in the build container; on the GPU box (where that tree does not exist) the git-ignored copy
    of the reference package that __graft_entry__.build() leaves under baseline/_ref/ and that ships with the
    working tree."""
    env = os.environ.get("QISKIT_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/qiskit"):
        return "/root/reference"
    return os.path.join(_REPO_ROOT, "baseline", "_ref")


REFERENCE_ROOT = _default_root()


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "qiskit", "providers", "basicaer"))


class _Missing:
    """Attribute bag whose members raise only when they are *used*."""

    def __init__(self, name):
        self._name = name

    def __call__(self, *a, **k):
        raise ImportError("stub for %s: not available in this environment" % self._name)

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        return _Missing(self._name + "." + item)


def _stub_module(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)

    def _getattr(item, _name=name):
        if item.startswith("__"):
            raise AttributeError(item)
        # classes are used as base classes / in isinstance checks at import time
        return type(item, (), {"__init__": lambda self, *a, **k: (_ for _ in ()).throw(
            ImportError("stub for %s.%s" % (_name, item)))})

    mod.__getattr__ = _getattr
    sys.modules[name] = mod
    return mod


def _install_stubs():
    import numpy as np

    if not hasattr(np, "product"):  # removed in numpy 2; used by quantum_info (operators/random.py:53)
        np.product = np.prod
    if not hasattr(np, "unicode_"):  # removed in numpy 2; used by result/counts handling (opflow state functions)
        np.unicode_ = np.str_
    if "retworkx" not in sys.modules:
        # pure-Python stand-in for the part of retworkx DAGCircuit uses (retworkx_shim.py), so that the
        # reference's transpile()/execute() can run here
        rx = _load_sibling("retworkx", "retworkx_shim.py")
        sys.modules["retworkx"] = rx
        _stub_module("retworkx.visualization")
        rx.visualization = sys.modules["retworkx.visualization"]
    if "ddt" not in sys.modules:
        try:
            importlib.import_module("ddt")
        except ImportError:
            sys.modules["ddt"] = _load_sibling("ddt", "ddt_shim.py")
    if "ply" not in sys.modules:
        ply = _stub_module("ply")
        ply.lex = _stub_module("ply.lex")
        ply.yacc = _stub_module("ply.yacc")
    if "fastjsonschema" not in sys.modules:
        fj = _stub_module("fastjsonschema", compile=lambda schema, *a, **k: (lambda data: data))
        exc = _stub_module("fastjsonschema.exceptions", JsonSchemaException=type(
            "JsonSchemaException", (Exception,), {}))
        fj.exceptions = exc
        fj.JsonSchemaException = exc.JsonSchemaException
    if "constraint" not in sys.modules:
        _stub_module("constraint")
    # the reference's expectation-value extension, if oracle/build_ref.py compiled it into oracle/_ref/
    real = _load_compiled_exp_value()
    if real is not None:
        sys.modules["qiskit.quantum_info.states.cython.exp_value"] = real
    # Cython extensions (not built in the read-only tree, none on the BasicAer path)
    for name in (
        "qiskit.quantum_info.states.cython.exp_value",
        "qiskit.transpiler.passes.routing.cython.stochastic_swap.utils",
        "qiskit.transpiler.passes.routing.cython.stochastic_swap.swap_trial",
    ):
        if name not in sys.modules:
            _stub_module(name)


def _load_sibling(name, filename):
    import importlib.util

    spec = importlib.util.spec_from_file_location(name, os.path.join(os.path.dirname(os.path.abspath(__file__)), filename))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _load_compiled_exp_value():
    import glob
    import importlib.machinery
    import importlib.util

    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cands = glob.glob(os.path.join(here, "_ref", "exp_value*.so"))
    if not cands:
        return None
    name = "qiskit.quantum_info.states.cython.exp_value"
    loader = importlib.machinery.ExtensionFileLoader(name, cands[0])
    spec = importlib.util.spec_from_file_location(name, cands[0], loader=loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod


def import_reference():
    """Return the reference ``qiskit`` package, imported from REFERENCE_ROOT."""
    if not reference_available():
        raise ImportError("reference tree not found at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # namespace-ish parent for the un-__init__'d cython dir
    qiskit = importlib.import_module("qiskit")
    if getattr(sys.modules.get("ply"), "__file__", None) is None:
        _install_qasm_reader(qiskit)
    return qiskit


def _install_qasm_reader(qiskit):
    """Without ply the reference cannot parse OpenQASM; route the two loaders to qasm2_reader.py."""
    reader = _load_sibling("_qsv_qasm2_reader", "qasm2_reader.py")
    qiskit.QuantumCircuit.from_qasm_file = staticmethod(lambda path: reader.circuit_from_qasm_file(qiskit, path))
    qiskit.QuantumCircuit.from_qasm_str = staticmethod(lambda text: reader.circuit_from_qasm_str(qiskit, text))
