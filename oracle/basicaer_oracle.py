"""CPU oracle: a numpy restatement of the reference's BasicAer statevector hot path.

TEST INFRASTRUCTURE ONLY.  This module is the *checker* for the CUDA path.  It may be
imported only by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs.  Nothing under ``qiskit_sdk_py_b200/`` imports it, and the product
path fails loudly (``QsvLibraryError``) when the CUDA library is missing -- there is no CPU
fallback.

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the unmodified reference
(``/root/reference``, Qiskit Terra 0.18.0, imported through ``oracle/ref_env/ref_import.py``)
on every fixture under ``tests/golden/`` and records its outputs; ``tests/test_oracle_golden.py``
checks this restatement against those recorded outputs (amplitudes <= 1e-13, counts and
memory lists identical).  When ``/root/reference`` is present the same test module also runs
the reference live.

The arithmetic the reference uses on this path is numpy's (not vendored; requirements.txt:4
``numpy>=1.17``, installed here 2.3.5): ``np.einsum`` for gate application, ``np.sum(np.abs()**2)``
for probabilities and the legacy ``np.random.RandomState`` (MT19937, frozen stream) for
``rand()``/``choice()``.  This restatement calls the same numpy primitives in the same order,
so it reproduces the reference to the last bit for the operations it covers; what is
restated is the *control flow and index conventions* around them.

Input format: the Qobj wire format (``QasmQobj.to_dict()``; qiskit/qobj/qasm_qobj.py:30-102,
203-330, 538-650), a plain dict -- no qiskit import is needed.

Every function cites the reference lines it follows (paths relative to /root/reference).
"""
import time
from collections import Counter

import numpy as np

MAX_QUBITS = 24  # qiskit/providers/basicaer/qasm_simulator.py:64

# qiskit/providers/basicaer/basicaertools.py:26
SINGLE_QUBIT_GATES = ("U", "u1", "u2", "u3", "rz", "sx", "x")


class OracleError(Exception):
    """Stands in for BasicAerError (qiskit/providers/basicaer/exceptions.py:20-30)."""


# --------------------------------------------------------------------------------------
# gate matrices
# --------------------------------------------------------------------------------------
def single_gate_matrix(name, params=None):
    """2x2 matrix of a basis gate.

    Follows qiskit/providers/basicaer/basicaertools.py:29-63, which instantiates the gate
    class and calls ``to_matrix()``; the closed forms are the ``__array__`` bodies:
    UGate circuit/library/standard_gates/u.py:98-110, U3Gate u3.py:98-110, U2Gate
    u2.py:85-96, U1Gate u1.py:121-124, RZGate rz.py:103-108, SXGate sx.py:100-102,
    XGate x.py:115-117, IGate i.py.  The expressions are written with the same numpy
    calls so the entries are bit-identical to the reference's.
    """
    params = [] if params is None else [float(p) for p in params]
    if name in ("U", "u3"):
        theta, phi, lam = params
        cos = np.cos(theta / 2)
        sin = np.sin(theta / 2)
        return np.array(
            [
                [cos, -np.exp(1j * lam) * sin],
                [np.exp(1j * phi) * sin, np.exp(1j * (phi + lam)) * cos],
            ],
            dtype=complex,
        )
    if name == "u2":
        phi, lam = params
        isqrt2 = 1 / np.sqrt(2)
        return np.array(
            [
                [isqrt2, -np.exp(1j * lam) * isqrt2],
                [np.exp(1j * phi) * isqrt2, np.exp(1j * (phi + lam)) * isqrt2],
            ],
            dtype=complex,
        )
    if name == "u1":
        (lam,) = params
        return np.array([[1, 0], [0, np.exp(1j * lam)]], dtype=complex)
    if name == "rz":
        (lam,) = params
        ilam2 = 0.5j * lam
        return np.array([[np.exp(-ilam2), 0], [0, np.exp(ilam2)]], dtype=complex)
    if name == "sx":
        return np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]], dtype=complex) / 2
    if name == "x":
        return np.array([[0, 1], [1, 0]], dtype=complex)
    if name == "id":
        return np.array([[1, 0], [0, 1]], dtype=complex)
    raise OracleError("Gate is not a valid basis gate for this simulator: %s" % name)


def cx_gate_matrix():
    """qiskit/providers/basicaer/basicaertools.py:70-72 (qubits passed as [control, target])."""
    return np.array([[1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0], [0, 1, 0, 0]], dtype=complex)


# --------------------------------------------------------------------------------------
# gate application
# --------------------------------------------------------------------------------------
def apply_unitary(state_tensor, gate, qubits):
    """psi <- (gate on ``qubits``) psi, out of place, with one np.einsum.

    Follows QasmSimulatorPy._add_unitary (qiskit/providers/basicaer/qasm_simulator.py:145-161)
    and the index convention of _einsum_matmul_index_helper (basicaertools.py:132-180):
    the state is a rank-n tensor whose axis ``n-1-q`` is qubit ``q``; the gate is reshaped to
    rank 2k with the k row axes first; row/column axis ``k-1-j`` belongs to ``qubits[j]``
    (``qubits[0]`` is the least significant bit of the matrix index).

    The reference spells the contraction with letters (which caps n+k at 26,
    basicaertools.py:158-159); here it is spelled with einsum's integer-sublist form, which
    lowers to the same C inner loop.
    """
    n = state_tensor.ndim
    k = len(qubits)
    gate_tensor = np.reshape(np.array(gate, dtype=complex), k * [2, 2])
    in_labels = list(range(n))
    out_labels = list(range(n))
    row_labels = [0] * k
    col_labels = [0] * k
    for j, q in enumerate(qubits):
        axis = n - 1 - q
        fresh = n + j
        col_labels[k - 1 - j] = in_labels[axis]
        row_labels[k - 1 - j] = fresh
        out_labels[axis] = fresh
    return np.einsum(
        gate_tensor,
        row_labels + col_labels,
        state_tensor,
        in_labels,
        out_labels,
        dtype=complex,
        casting="no",
    )


def apply_unitary_rows(unitary_tensor, gate, qubits, n):
    """U <- (gate (x) I) U on the row-index bits of a rank-2n tensor.

    Follows UnitarySimulatorPy._add_unitary (qiskit/providers/basicaer/unitary_simulator.py:127-141)
    with einsum_matmul_index (basicaertools.py:75-103): the first n axes are the row index
    (axis ``n-1-q`` = qubit q), the last n axes the column index, which passes through.
    """
    k = len(qubits)
    gate_tensor = np.reshape(np.array(gate, dtype=complex), k * [2, 2])
    in_labels = list(range(2 * n))
    out_labels = list(range(2 * n))
    row_labels = [0] * k
    col_labels = [0] * k
    for j, q in enumerate(qubits):
        axis = n - 1 - q
        fresh = 2 * n + j
        col_labels[k - 1 - j] = in_labels[axis]
        row_labels[k - 1 - j] = fresh
        out_labels[axis] = fresh
    return np.einsum(
        gate_tensor,
        row_labels + col_labels,
        unitary_tensor,
        in_labels,
        out_labels,
        dtype=complex,
        casting="no",
    )


# --------------------------------------------------------------------------------------
# qasm / statevector simulator
# --------------------------------------------------------------------------------------
class OracleQasmSimulator:
    """Restates QasmSimulatorPy / StatevectorSimulatorPy over Qobj dicts.

    ``show_final_state=True`` gives StatevectorSimulatorPy
    (qiskit/providers/basicaer/statevector_simulator.py:85,111-122: SHOW_FINAL_STATE and
    shots forced to 1).
    """

    def __init__(self, show_final_state=False, max_qubits=MAX_QUBITS):
        self.show_final_state = show_final_state
        self.max_qubits = max_qubits
        self.name = "statevector_simulator" if show_final_state else "qasm_simulator"
        self._local_random = np.random.RandomState()  # qasm_simulator.py:119
        self._statevector = 0
        self._classical_memory = 0
        self._classical_register = 0
        self._n = 0
        self._initial_statevector = None
        self._chop_threshold = 1e-15  # qasm_simulator.py:104,139
        self._sample_measure = False
        self._shots = 0
        self._memory = False
        self.gate_passes = 0  # instrumentation for the CPU baseline (not in the reference)

    # -- options -------------------------------------------------------------------
    def _set_options(self, qobj_config, backend_options):
        """qasm_simulator.py:289-317."""
        self._initial_statevector = None
        self._chop_threshold = 1e-15
        backend_options = dict(backend_options or {})
        if backend_options.get("backend_options"):
            backend_options = backend_options["backend_options"]
        if "initial_statevector" in backend_options:
            self._initial_statevector = np.array(backend_options["initial_statevector"], dtype=complex)
        elif "initial_statevector" in qobj_config:
            self._initial_statevector = np.array(qobj_config["initial_statevector"], dtype=complex)
        if self._initial_statevector is not None:
            norm = np.linalg.norm(self._initial_statevector)
            if round(norm, 12) != 1:
                raise OracleError(
                    "initial statevector is not normalized: " + "norm {} != 1".format(norm)
                )
        if "chop_threshold" in backend_options:
            self._chop_threshold = backend_options["chop_threshold"]
        elif "chop_threshold" in qobj_config:
            self._chop_threshold = qobj_config["chop_threshold"]

    def _validate(self, qobj):
        """qasm_simulator.py:655-664 (the logger warnings at :665-675 are not restated)."""
        n_qubits = qobj["config"]["n_qubits"]
        if n_qubits > self.max_qubits:
            raise OracleError(
                "Number of qubits {} ".format(n_qubits)
                + "is greater than maximum ({}) ".format(self.max_qubits)
                + 'for "{}".'.format(self.name)
            )

    # -- state ---------------------------------------------------------------------
    def _initialize_statevector(self):
        """qasm_simulator.py:319-328."""
        if self._initial_statevector is None:
            self._statevector = np.zeros(2 ** self._n, dtype=complex)
            self._statevector[0] = 1
        else:
            self._statevector = self._initial_statevector.copy()
        self._statevector = np.reshape(self._statevector, self._n * [2])

    def _add_unitary(self, gate, qubits):
        """qasm_simulator.py:145-161."""
        self._statevector = apply_unitary(self._statevector, gate, qubits)
        self.gate_passes += 1

    def _get_measure_outcome(self, qubit):
        """qasm_simulator.py:163-182: (p0, p1) by np.sum over all other axes, one rand()."""
        axis = list(range(self._n))
        axis.remove(self._n - 1 - qubit)
        probabilities = np.sum(np.abs(self._statevector) ** 2, axis=tuple(axis))
        random_number = self._local_random.rand()
        if random_number < probabilities[0]:
            return "0", probabilities[0]
        return "1", probabilities[1]

    def _add_qasm_measure(self, qubit, cmembit, cregbit=None):
        """qasm_simulator.py:227-253."""
        outcome, probability = self._get_measure_outcome(qubit)
        membit = 1 << cmembit
        self._classical_memory = (self._classical_memory & (~membit)) | (int(outcome) << cmembit)
        if cregbit is not None:
            regbit = 1 << cregbit
            self._classical_register = (self._classical_register & (~regbit)) | (
                int(outcome) << cregbit
            )
        if outcome == "0":
            update_diag = [[1 / np.sqrt(probability), 0], [0, 0]]
        else:
            update_diag = [[0, 0], [0, 1 / np.sqrt(probability)]]
        self._add_unitary(update_diag, [qubit])

    def _add_qasm_reset(self, qubit):
        """qasm_simulator.py:255-273."""
        outcome, probability = self._get_measure_outcome(qubit)
        if outcome == "0":
            update = [[1 / np.sqrt(probability), 0], [0, 0]]
        else:
            update = [[0, 1 / np.sqrt(probability)], [0, 0]]
        self._add_unitary(update, [qubit])

    def _add_sample_measure(self, measure_params, num_samples):
        """qasm_simulator.py:184-225: marginal over the sorted measured qubits, then
        RandomState.choice (= cumsum, normalise, searchsorted(random_sample, 'right'))."""
        measured_qubits = sorted({qubit for qubit, _ in measure_params})
        num_measured = len(measured_qubits)
        axis = list(range(self._n))
        for qubit in reversed(measured_qubits):
            axis.remove(self._n - 1 - qubit)
        probabilities = np.reshape(
            np.sum(np.abs(self._statevector) ** 2, axis=tuple(axis)), 2 ** num_measured
        )
        samples = self._local_random.choice(range(2 ** num_measured), num_samples, p=probabilities)
        memory = []
        for sample in samples:
            classical_memory = self._classical_memory
            for qubit, cmembit in measure_params:
                pos = measured_qubits.index(qubit)
                qubit_outcome = int((int(sample) & (1 << pos)) >> pos)
                membit = 1 << cmembit
                classical_memory = (classical_memory & (~membit)) | (qubit_outcome << cmembit)
            memory.append(hex(classical_memory))
        return memory

    def _get_statevector(self):
        """qasm_simulator.py:330-334."""
        vec = np.reshape(self._statevector, 2 ** self._n)
        vec[abs(vec) < self._chop_threshold] = 0.0
        return vec

    def _validate_measure_sampling(self, experiment):
        """qasm_simulator.py:336-374."""
        if self._shots <= 1:
            self._sample_measure = False
            return
        if "allows_measure_sampling" in experiment.get("config", {}):
            self._sample_measure = experiment["config"]["allows_measure_sampling"]
            return
        measure_flag = False
        for instruction in experiment["instructions"]:
            name = instruction["name"]
            if name == "reset":
                self._sample_measure = False
                return
            if measure_flag:
                if name not in ["measure", "barrier", "id", "u0"]:
                    self._sample_measure = False
                    return
            elif name == "measure":
                measure_flag = True
        self._sample_measure = True

    # -- run -----------------------------------------------------------------------
    def run(self, qobj, **backend_options):
        """qasm_simulator.py:376-457 for a Qobj dict; returns the result dict that the
        reference feeds to Result.from_dict."""
        qobj_config = dict(qobj["config"])
        if self.show_final_state:
            # statevector_simulator.py:111-122 forces shots to 1
            qobj_config["shots"] = 1
            for experiment in qobj["experiments"]:
                if "shots" in experiment.get("config", {}):
                    experiment["config"]["shots"] = 1
        self._set_options(qobj_config, backend_options)
        self._validate(qobj)
        self._shots = qobj_config["shots"]
        self._memory = qobj_config.get("memory", False)
        self._qobj_config = qobj_config
        start = time.time()
        results = [self.run_experiment(exp) for exp in qobj["experiments"]]
        end = time.time()
        return {
            "backend_name": self.name,
            "backend_version": "2.1.0" if not self.show_final_state else "1.1.0",
            "qobj_id": qobj.get("qobj_id"),
            "job_id": "oracle",
            "results": results,
            "status": "COMPLETED",
            "success": True,
            "time_taken": end - start,
            "header": qobj.get("header", {}),
        }

    def run_experiment(self, experiment):
        """qasm_simulator.py:459-653."""
        start = time.time()
        exp_config = experiment.get("config", {})
        header = experiment.get("header", {})
        self._n = exp_config["n_qubits"]
        number_of_cmembits = exp_config["memory_slots"]
        self._statevector = 0
        self._classical_memory = 0
        self._classical_register = 0
        self._sample_measure = False
        global_phase = header.get("global_phase", 0.0)
        if self._initial_statevector is not None:  # :275-287
            length = len(self._initial_statevector)
            required_dim = 2 ** self._n
            if length != required_dim:
                raise OracleError(
                    "initial statevector is incorrect length: "
                    + "{} != {}".format(length, required_dim)
                )
        if "seed_simulator" in exp_config:  # :495-502
            seed_simulator = exp_config["seed_simulator"]
        elif "seed_simulator" in self._qobj_config:
            seed_simulator = self._qobj_config["seed_simulator"]
        else:
            seed_simulator = np.random.randint(2147483647, dtype="int32")
        self._local_random.seed(seed=seed_simulator)
        self._validate_measure_sampling(experiment)

        memory = []
        if self._sample_measure:
            shots = 1
            measure_sample_ops = []
        else:
            shots = self._shots
        for _ in range(shots):
            self._initialize_statevector()
            self._statevector *= np.exp(1j * global_phase)  # :522
            self._classical_memory = 0
            self._classical_register = 0
            for operation in experiment["instructions"]:
                conditional = operation.get("conditional", None)
                if isinstance(conditional, int):  # :527-531
                    if not (self._classical_register >> conditional) & 1:
                        continue
                elif conditional is not None:  # legacy mask/val form, :532-540
                    mask = int(conditional["mask"], 16)
                    if mask > 0:
                        value = self._classical_memory & mask
                        while (mask & 0x1) == 0:
                            mask >>= 1
                            value >>= 1
                        if value != int(conditional["val"], 16):
                            continue
                name = operation["name"]
                if name == "unitary":
                    self._add_unitary(operation["params"][0], operation["qubits"])
                elif name in SINGLE_QUBIT_GATES:
                    gate = single_gate_matrix(name, operation.get("params", None))
                    self._add_unitary(gate, [operation["qubits"][0]])
                elif name in ("id", "u0"):
                    pass
                elif name in ("CX", "cx"):
                    self._add_unitary(cx_gate_matrix(), operation["qubits"][:2])
                elif name == "reset":
                    self._add_qasm_reset(operation["qubits"][0])
                elif name == "barrier":
                    pass
                elif name == "measure":
                    qubit = operation["qubits"][0]
                    cmembit = operation["memory"][0]
                    cregbit = operation["register"][0] if "register" in operation else None
                    if self._sample_measure:
                        measure_sample_ops.append((qubit, cmembit))
                    else:
                        self._add_qasm_measure(qubit, cmembit, cregbit)
                elif name == "bfunc":  # :580-614
                    mask = int(operation["mask"], 16)
                    relation = operation["relation"]
                    val = int(operation["val"], 16)
                    cregbit = operation["register"]
                    cmembit = operation.get("memory", None)
                    compared = (self._classical_register & mask) - val
                    if relation == "==":
                        outcome = compared == 0
                    elif relation == "!=":
                        outcome = compared != 0
                    elif relation == "<":
                        outcome = compared < 0
                    elif relation == "<=":
                        outcome = compared <= 0
                    elif relation == ">":
                        outcome = compared > 0
                    elif relation == ">=":
                        outcome = compared >= 0
                    else:
                        raise OracleError("Invalid boolean function relation.")
                    regbit = 1 << cregbit
                    self._classical_register = (self._classical_register & (~regbit)) | (
                        int(outcome) << cregbit
                    )
                    if cmembit is not None:
                        membit = 1 << cmembit
                        self._classical_memory = (self._classical_memory & (~membit)) | (
                            int(outcome) << cmembit
                        )
                else:
                    err_msg = '{0} encountered unrecognized operation "{1}"'
                    raise OracleError(err_msg.format(self.name, name))
            if number_of_cmembits > 0:  # :621-628
                if self._sample_measure:
                    memory = self._add_sample_measure(measure_sample_ops, self._shots)
                else:
                    memory.append(hex(self._classical_memory))

        data = {"counts": dict(Counter(memory))}
        if self._memory:
            data["memory"] = memory
        if self.show_final_state:
            data["statevector"] = self._get_statevector()
            if not data["counts"]:
                data.pop("counts")
            if "memory" in data and not data["memory"]:
                data.pop("memory")
        end = time.time()
        return {
            "name": header.get("name"),
            "seed_simulator": seed_simulator,
            "shots": self._shots,
            "data": data,
            "status": "DONE",
            "success": True,
            "time_taken": end - start,
            "header": header,
        }


# --------------------------------------------------------------------------------------
# unitary simulator
# --------------------------------------------------------------------------------------
class OracleUnitarySimulator:
    """Restates UnitarySimulatorPy (qiskit/providers/basicaer/unitary_simulator.py:57-401)."""

    name = "unitary_simulator"

    def __init__(self, max_qubits=MAX_QUBITS):
        self.max_qubits = max_qubits
        self._initial_unitary = None
        self._chop_threshold = 1e-15
        self._global_phase = 0.0

    def _set_options(self, qobj_config, backend_options):
        """unitary_simulator.py:143-188."""
        self._initial_unitary = None
        self._chop_threshold = 1e-15
        backend_options = dict(backend_options or {})
        if backend_options.get("backend_options"):
            backend_options = backend_options["backend_options"]
        if "initial_unitary" in backend_options:
            self._initial_unitary = np.array(backend_options["initial_unitary"], dtype=complex)
        elif "initial_unitary" in qobj_config:
            self._initial_unitary = np.array(qobj_config["initial_unitary"], dtype=complex)
        if self._initial_unitary is not None:
            shape = np.shape(self._initial_unitary)
            if len(shape) != 2 or shape[0] != shape[1]:
                raise OracleError("initial unitary is not a square matrix")
            iden = np.eye(len(self._initial_unitary))
            u_dagger_u = np.dot(self._initial_unitary.T.conj(), self._initial_unitary)
            norm = np.linalg.norm(u_dagger_u - iden)
            if round(norm, 10) != 0:
                raise OracleError("initial unitary is not unitary")
        if "chop_threshold" in backend_options:
            self._chop_threshold = backend_options["chop_threshold"]
        elif "chop_threshold" in qobj_config:
            self._chop_threshold = qobj_config["chop_threshold"]

    def run(self, qobj, **backend_options):
        """unitary_simulator.py:209-292."""
        self._set_options(qobj["config"], backend_options)
        n_qubits = qobj["config"]["n_qubits"]
        if n_qubits > self.max_qubits:
            raise OracleError(
                "Number of qubits {} ".format(n_qubits)
                + "is greater than maximum ({}) ".format(self.max_qubits)
                + 'for "{}".'.format(self.name)
            )
        start = time.time()
        results = [self.run_experiment(exp) for exp in qobj["experiments"]]
        end = time.time()
        return {
            "backend_name": self.name,
            "backend_version": "1.1.0",
            "qobj_id": qobj.get("qobj_id"),
            "job_id": "oracle",
            "results": results,
            "status": "COMPLETED",
            "success": True,
            "time_taken": end - start,
            "header": qobj.get("header", {}),
        }

    def run_experiment(self, experiment):
        """unitary_simulator.py:294-366 (+ _validate :368-401 for measure/reset)."""
        start = time.time()
        header = experiment.get("header", {})
        n = experiment["config"]["n_qubits"]
        self._global_phase = header.get("global_phase", 0.0)
        for operation in experiment["instructions"]:
            if operation["name"] in ["measure", "reset"]:
                raise OracleError(
                    'Unsupported "%s" instruction "%s" ' + 'in circuit "%s" ',
                    self.name,
                    operation["name"],
                    header.get("name"),
                )
        if self._initial_unitary is not None:  # :182-188, :190-199
            required_shape = (2 ** n, 2 ** n)
            if np.shape(self._initial_unitary) != required_shape:
                raise OracleError(
                    "initial unitary is incorrect shape: "
                    + "{} != 2 ** {}".format(np.shape(self._initial_unitary), required_shape)
                )
            unitary = self._initial_unitary.copy()
        else:
            unitary = np.eye(2 ** n, dtype=complex)
        unitary = np.reshape(unitary, n * [2, 2])
        for operation in experiment["instructions"]:
            name = operation["name"]
            if name == "unitary":
                unitary = apply_unitary_rows(unitary, operation["params"][0], operation["qubits"], n)
            elif name in ("id", "u0"):
                pass
            elif name in SINGLE_QUBIT_GATES:
                gate = single_gate_matrix(name, operation.get("params", None))
                unitary = apply_unitary_rows(unitary, gate, [operation["qubits"][0]], n)
            elif name in ("CX", "cx"):
                unitary = apply_unitary_rows(unitary, cx_gate_matrix(), operation["qubits"][:2], n)
            elif name == "barrier":
                pass
            else:
                err_msg = '{0} encountered unrecognized operation "{1}"'
                raise OracleError(err_msg.format(self.name, name))
        # _get_unitary, :201-207
        unitary = np.reshape(unitary, 2 * [2 ** n])
        if self._global_phase:
            unitary *= np.exp(1j * float(self._global_phase))
        unitary[abs(unitary) < self._chop_threshold] = 0.0
        end = time.time()
        return {
            "name": header.get("name"),
            "shots": 1,
            "data": {"unitary": unitary},
            "status": "DONE",
            "success": True,
            "time_taken": end - start,
            "header": header,
        }


# --------------------------------------------------------------------------------------
# Pauli expectation values (SURVEY 8(f) rank 4 -- the reference's only native code near the path)
# --------------------------------------------------------------------------------------
def _parity(values):
    v = values.astype(np.uint64)
    for shift in (32, 16, 8, 4, 2, 1):
        v ^= v >> np.uint64(shift)
    return (v & np.uint64(1)).astype(bool)


def expval_pauli(data, label):
    """<psi|P|psi> for a Pauli label ("XIZY": leftmost character = highest qubit).

    Restates Statevector._expectation_value_pauli (qiskit/quantum_info/states/statevector.py:377-408)
    with the Cython loops expval_pauli_no_x / expval_pauli_with_x
    (qiskit/quantum_info/states/cython/exp_value.pyx:37-49, :52-94) vectorised in numpy.  For a Pauli
    built from a plain label the group phase is 0, so only the (-i)^(#Y) factor inside the loop remains.
    Note the reference's quirk for the identity: it returns the norm, not the squared norm (:397-398)."""
    data = np.asarray(data, dtype=complex).reshape(-1)
    n = len(label)
    assert data.size == 2 ** n
    z_mask = x_mask = n_y = 0
    for q, c in enumerate(reversed(label.upper())):
        if c in "ZY":
            z_mask |= 1 << q
        if c in "XY":
            x_mask |= 1 << q
        n_y += c == "Y"
    if x_mask + z_mask == 0:
        return float(np.linalg.norm(data))
    idx = np.arange(data.size, dtype=np.uint64)
    if x_mask == 0:
        p = data.real * data.real + data.imag * data.imag
        sign = np.where(_parity(idx & np.uint64(z_mask)), -1.0, 1.0)
        return float(np.sum(sign * p))
    phase = (-1j) ** n_y
    x_max = x_mask.bit_length() - 1
    i = np.arange(data.size // 2, dtype=np.uint64)
    mask_u = np.uint64(~((2 << x_max) - 1) & (2 ** 64 - 1))
    mask_l = np.uint64((1 << x_max) - 1)
    i0 = ((i << np.uint64(1)) & mask_u) | (i & mask_l)
    i1 = i0 ^ np.uint64(x_mask)
    d0, d1 = data[i0.astype(np.int64)], data[i1.astype(np.int64)]
    v0 = (phase * (d1 * np.conj(d0))).real
    v1 = (phase * (d0 * np.conj(d1))).real
    s0 = np.where(_parity(i0 & np.uint64(z_mask)), -1.0, 1.0)
    s1 = np.where(_parity(i1 & np.uint64(z_mask)), -1.0, 1.0)
    return float(np.sum(s0 * v0) + np.sum(s1 * v1))
