"""Run the reference's OWN BasicAer test files, unmodified and read where they lie, against the drop-in
backends ("test/python/basicaer/* run against the new backend").

``qiskit.BasicAer`` and the three ``*SimulatorPy`` classes are swapped for the B200 provider/backends for
the duration of the run; the device is tests/fake_state.FakeState, so this covers the whole host side
(provider, job, option handling, lowering, fusion planner, control flow, Result) through ``execute()``,
``transpile()``, ``assemble()`` and the OpenQASM loader exactly as the reference's tests drive it.
Skipped where the reference tree is not available (the GPU box)."""
import glob
import importlib.util
import os
import sys
import unittest
import warnings

import pytest

REF = "/root/reference"
TEST_DIR = os.path.join(REF, "test", "python", "basicaer")
pytestmark = pytest.mark.skipif(not os.path.isdir(TEST_DIR), reason="reference tests not present")


@pytest.fixture(scope="module")
def qk():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "ref_env"))
    import ref_import

    ref_import.import_reference()
    warnings.simplefilter("ignore")
    import qiskit

    return qiskit


def _drop_in_classes():
    from fake_state import FakeState
    from qiskit_sdk_py_b200 import provider as prov

    kwargs = {"state_factory": FakeState, "max_qubits": 24}

    def with_fake_device(cls):
        class _Backend(cls):
            def __init__(self, configuration=None, provider=None, **fields):
                super().__init__(configuration=configuration, provider=provider, simulator_kwargs=kwargs, **fields)

        _Backend.__name__ = cls.__name__
        return _Backend

    class _Provider(prov.B200Provider):
        def __init__(self):
            super().__init__(simulator_kwargs=kwargs)

    return (_Provider, with_fake_device(prov.QasmSimulatorB200Backend),
            with_fake_device(prov.StatevectorSimulatorB200Backend),
            with_fake_device(prov.UnitarySimulatorB200Backend))


def _load(path, tag):
    name = "_refsuite_%s_%s" % (tag, os.path.basename(path)[:-3])
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _run(tag):
    suite = unittest.TestSuite()
    loader = unittest.TestLoader()
    for path in sorted(glob.glob(os.path.join(TEST_DIR, "test_*.py"))):
        suite.addTests(loader.loadTestsFromModule(_load(path, tag)))
    result = unittest.TestResult()
    suite.run(result)
    return result


def _counting(orig, runs):
    def run(self, *args, **kwargs):
        runs.append(type(self).__name__)
        return orig(self, *args, **kwargs)

    return run


def _describe(result):
    return "\n".join("%s\n%s" % (t.id(), tb) for t, tb in result.failures + result.errors)


def test_reference_basicaer_tests_pass_on_the_reference(qk):
    """Control: with the shims (retworkx, OpenQASM reader) the reference's tests pass on the reference."""
    result = _run("ref")
    assert result.testsRun >= 30
    assert result.wasSuccessful(), _describe(result)


def test_reference_basicaer_tests_pass_on_drop_in(qk):
    import qiskit.providers.basicaer as basicaer

    from qiskit_sdk_py_b200 import simulator

    provider_cls, qasm_cls, sv_cls, unitary_cls = _drop_in_classes()
    runs = []
    with pytest.MonkeyPatch.context() as mp:
        for cls in (simulator.QasmSimulatorB200, simulator.StatevectorSimulatorB200, simulator.UnitarySimulatorB200):
            if "run" in vars(cls):
                mp.setattr(cls, "run", _counting(cls.run, runs))
        instance = provider_cls()
        mp.setattr(basicaer, "BasicAerProvider", provider_cls)
        mp.setattr(basicaer, "BasicAer", instance)
        mp.setattr(qk, "BasicAer", instance)
        mp.setattr(basicaer, "QasmSimulatorPy", qasm_cls)
        mp.setattr(basicaer, "StatevectorSimulatorPy", sv_cls)
        mp.setattr(basicaer, "UnitarySimulatorPy", unitary_cls)
        result = _run("b200")
    assert result.testsRun >= 30
    assert result.wasSuccessful(), _describe(result)
    assert not result.skipped, result.skipped
    assert len(runs) >= 40, "the reference's tests did not go through the drop-in simulators"
