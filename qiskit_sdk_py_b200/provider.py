"""Qiskit Terra 0.18 provider plugin: the B200 simulators behind the reference's own plugin API.

Importing this module requires ``qiskit`` (Terra 0.18).  It defines

    B200Provider(ProviderV1)              mirror of BasicAerProvider   (basicaerprovider.py:34-128)
    QasmSimulatorB200Backend(BackendV1)   mirror of QasmSimulatorPy    (qasm_simulator.py:56-143,376-457)
    StatevectorSimulatorB200Backend       mirror of StatevectorSimulatorPy (statevector_simulator.py:36-122)
    UnitarySimulatorB200Backend           mirror of UnitarySimulatorPy (unitary_simulator.py:57-125,209-262)
    B200Job(JobV1)                        mirror of BasicAerJob        (basicaerjob.py:24-68)

with the same backend names, configuration fields, default options, run() signature
(``QuantumCircuit`` | list | ``QasmQobj``), warnings and ``Result`` layout, so that

    from qiskit_sdk_py_b200.provider import B200Aer
    backend = B200Aer.get_backend("qasm_simulator")
    counts = execute(circuit, backend, shots=1024, seed_simulator=42).result().get_counts()

is a drop-in for ``BasicAer``.  The arithmetic runs in ``simulator.py`` -> ``engine.py`` -> libqsv (CUDA);
this file only converts ``QasmQobj.to_dict()`` in and ``Result.from_dict()`` out.
"""
import uuid
import warnings
from collections import OrderedDict

from qiskit.circuit import QuantumCircuit
from qiskit.exceptions import QiskitError
from qiskit.providers import JobStatus
from qiskit.providers.backend import BackendV1
from qiskit.providers.basicaer.exceptions import BasicAerError
from qiskit.providers.exceptions import QiskitBackendNotFoundError
from qiskit.providers.job import JobV1
from qiskit.providers.models import QasmBackendConfiguration
from qiskit.providers.options import Options
from qiskit.providers.provider import ProviderV1
from qiskit.providers.providerutils import filter_backends, resolve_backend_name
from qiskit.result import Result

from . import simulator


class B200Job(JobV1):
    """Synchronous job: the result exists when run() returns (basicaerjob.py:24-68)."""

    _async = False

    def __init__(self, backend, job_id, result):
        super().__init__(backend, job_id)
        self._result = result

    def submit(self):
        return

    def result(self, timeout=None):  # pylint: disable=arguments-differ
        if timeout is not None:
            warnings.warn(
                "The timeout kwarg doesn't have any meaning with "
                "BasicAer because execution is synchronous and the "
                "result already exists when run() returns.",
                UserWarning,
            )
        return self._result

    def status(self):
        return JobStatus.DONE

    def backend(self):
        return self._backend


class _B200Backend(BackendV1):
    """Shared run() plumbing (qasm_simulator.py:376-424 / unitary_simulator.py:209-262)."""

    SIMULATOR_CLS = None

    def __init__(self, configuration=None, provider=None, simulator_kwargs=None, simulator_instance=None, **fields):
        self._sim = simulator_instance or self.SIMULATOR_CLS(**(simulator_kwargs or {}))
        super().__init__(
            configuration=(configuration or
                           QasmBackendConfiguration.from_dict(self._sim.configuration_dict())),
            provider=provider,
            **fields,
        )

    def run(self, qobj, **backend_options):
        """Run a circuit, a list of circuits or a (deprecated) Qobj; returns a job whose result()
        is a ``qiskit.result.Result``."""
        if isinstance(qobj, (QuantumCircuit, list)):
            from qiskit.compiler import assemble

            out_options = {}
            for key in backend_options:
                if not hasattr(self.options, key):
                    warnings.warn(
                        "Option %s is not used by this backend" % key, UserWarning, stacklevel=2
                    )
                else:
                    out_options[key] = backend_options[key]
            qobj = assemble(qobj, self, **out_options)
        else:
            warnings.warn(
                "Using a qobj for run() is deprecated and will be removed in a future release.",
                PendingDeprecationWarning,
                stacklevel=2,
            )
        job_id = str(uuid.uuid4())
        qobj_dict = qobj.to_dict()
        # enums (MeasLevel, ...) -> plain values, as they appear on the wire
        for key, val in list(qobj_dict["config"].items()):
            if hasattr(val, "value") and not isinstance(val, (int, float, str, list, dict)):
                qobj_dict["config"][key] = val.value
        # options set with set_options() persist across runs, as in the reference (qasm_simulator.py:292-293)
        self._sim.option_defaults = {key: getattr(self.options, key) for key in
                                     ("initial_statevector", "initial_unitary", "chop_threshold")
                                     if hasattr(self.options, key)}
        try:
            result = self._sim.run(qobj_dict, **backend_options)
        except simulator.BasicAerError as exc:
            raise BasicAerError(exc.message) from exc
        result["job_id"] = job_id
        result["backend_name"] = self.name()
        result["backend_version"] = self._configuration.backend_version
        return B200Job(self, job_id, Result.from_dict(result))


class QasmSimulatorB200Backend(_B200Backend):
    """B200 qasm simulator (drop-in for QasmSimulatorPy)."""

    SIMULATOR_CLS = simulator.QasmSimulatorB200

    @classmethod
    def _default_options(cls):
        return Options(**simulator.QasmSimulatorB200.DEFAULT_OPTIONS)


class StatevectorSimulatorB200Backend(_B200Backend):
    """B200 statevector simulator (drop-in for StatevectorSimulatorPy)."""

    SIMULATOR_CLS = simulator.StatevectorSimulatorB200

    @classmethod
    def _default_options(cls):
        return Options(**simulator.QasmSimulatorB200.DEFAULT_OPTIONS)


class UnitarySimulatorB200Backend(_B200Backend):
    """B200 unitary simulator (drop-in for UnitarySimulatorPy)."""

    SIMULATOR_CLS = simulator.UnitarySimulatorB200

    @classmethod
    def _default_options(cls):
        return Options(**simulator.UnitarySimulatorB200.DEFAULT_OPTIONS)


SIMULATORS = [QasmSimulatorB200Backend, StatevectorSimulatorB200Backend, UnitarySimulatorB200Backend]


class B200Provider(ProviderV1):
    """Provider for the B200 backends; same lookup rules as BasicAerProvider."""

    def __init__(self, simulator_kwargs=None, distributed=False, group=None, distributed_kwargs=None):
        """``distributed=True`` (inside a ``torch.distributed`` job, one process per GPU): the three
        backends keep their state sharded over the ranks of ``group`` (dist.distributed_simulator);
        every rank makes the same calls and gets the same Result, and ``n_qubits`` grows by log2(world) (the
        unitary backend's 2n-qubit vector is sharded the same way: log2(world) / 2 more qubits)."""
        super().__init__()
        self._simulator_kwargs = simulator_kwargs
        self._distributed = bool(distributed)
        self._group = group
        self._distributed_kwargs = dict(distributed_kwargs or {})
        self._backends = self._verify_backends()

    def get_backend(self, name=None, **kwargs):
        backends = self._backends.values()
        if name:
            try:
                name = resolve_backend_name(name, backends, self._deprecated_backend_names(), {})
            except LookupError as ex:
                raise QiskitBackendNotFoundError(
                    f"The '{name}' backend is not installed in your system."
                ) from ex
        return super().get_backend(name=name, **kwargs)

    def backends(self, name=None, filters=None, **kwargs):  # pylint: disable=arguments-differ
        backends = self._backends.values()
        if name:
            try:
                resolved_name = resolve_backend_name(name, backends, self._deprecated_backend_names(), {})
                backends = [backend for backend in backends if backend.name() == resolved_name]
            except LookupError:
                return []
        return filter_backends(backends, filters=filters, **kwargs)

    @staticmethod
    def _deprecated_backend_names():
        return {
            "qasm_simulator_py": "qasm_simulator",
            "statevector_simulator_py": "statevector_simulator",
            "unitary_simulator_py": "unitary_simulator",
            "local_qasm_simulator_py": "qasm_simulator",
            "local_statevector_simulator_py": "statevector_simulator",
            "local_unitary_simulator_py": "unitary_simulator",
            "local_unitary_simulator": "unitary_simulator",
        }

    def _verify_backends(self):
        ret = OrderedDict()
        for backend_cls in SIMULATORS:
            try:
                sim = None
                if self._distributed:
                    from . import dist

                    sim = dist.distributed_simulator(backend_cls.SIMULATOR_CLS.NAME, group=self._group,
                                                     **self._distributed_kwargs)
                backend_instance = backend_cls(provider=self, simulator_kwargs=self._simulator_kwargs,
                                               simulator_instance=sim)
            except Exception as err:
                raise QiskitError(f"Backend {backend_cls} could not be instantiated: {err}") from err
            ret[backend_instance.name()] = backend_instance
        return ret

    def __str__(self):
        return "B200Aer"


def _lazy_provider():
    return B200Provider()


class _LazyProvider:
    """``B200Aer`` is created on first use so that importing the module never touches CUDA (Terra forks
    worker processes in parallel_map, tools/parallel.py:147; no CUDA context may exist before that)."""

    _instance = None

    def __getattr__(self, item):
        if _LazyProvider._instance is None:
            _LazyProvider._instance = _lazy_provider()
        return getattr(_LazyProvider._instance, item)

    def __str__(self):
        return "B200Aer"


B200Aer = _LazyProvider()
