"""Qiskit-free mirror of the BasicAer simulators over the Qobj wire format.

``QasmSimulatorB200`` / ``StatevectorSimulatorB200`` / ``UnitarySimulatorB200`` take the dict form of a
``QasmQobj`` (``qobj.to_dict()``; qiskit/qobj/qasm_qobj.py) and return the dict that the reference
feeds to ``Result.from_dict`` -- same keys, same hex-keyed ``counts`` in first-occurrence order, same
``memory`` list, ``statevector`` / ``unitary`` as complex128 ndarrays, same option precedence, seeds
and error messages as

    qiskit/providers/basicaer/qasm_simulator.py         (run :376-424, _run_job :426-457,
                                                          run_experiment :459-653, _validate :655-675)
    qiskit/providers/basicaer/statevector_simulator.py  (:85, :100-122)
    qiskit/providers/basicaer/unitary_simulator.py      (:209-401)

The host keeps what the reference keeps on the host -- the instruction dispatch, classical
memory/register ints, conditionals, ``bfunc`` and the numpy ``RandomState`` stream -- and sends every
operation on the 2^n amplitudes to the GPU through ``engine.DeviceState`` (libqsv).  The
``qiskit_sdk_py_b200.provider`` module wraps these classes in ``BackendV1`` subclasses when qiskit is
importable.  There is no CPU path: ``state_factory`` defaults to the CUDA engine and construction
fails without a GPU.
"""
import logging
import time
import uuid
from collections import Counter
from math import log2

import numpy as np

from . import lower
from .engine import DeviceState

logger = logging.getLogger(__name__)

BASIS_GATES = ["u1", "u2", "u3", "rz", "sx", "x", "cx", "id", "unitary"]
HBM_RESERVE_BYTES = 8 << 30  # workspace / staging head-room kept free when sizing n_qubits


class BasicAerError(Exception):
    """Mirror of qiskit.providers.basicaer.exceptions.BasicAerError (exceptions.py:20-30); the
    provider adapter re-raises it as the reference's class."""

    def __init__(self, *message):
        super().__init__(" ".join(str(m) for m in message))
        self.message = " ".join(str(m) for m in message)

    def __str__(self):
        return repr(self.message)


def max_qubits_for_device(bytes_per_amp=16, default=33):
    """Largest n with 16*2^n bytes fitting the current GPU (33 on a 180 GB B200); the reference's
    analogue is MAX_QUBITS_MEMORY from host RAM (qasm_simulator.py:59,64).  Never creates a CUDA
    context (Terra may fork worker processes after the backend object exists): before CUDA is
    initialised the B200 value is returned."""
    try:
        import torch

        if torch.cuda.is_available() and torch.cuda.is_initialized():
            total = torch.cuda.get_device_properties(torch.cuda.current_device()).total_memory
            return int(log2(max(1, total - HBM_RESERVE_BYTES) / bytes_per_amp))
    except Exception:  # pragma: no cover
        pass
    return default


def _touch_cuda():
    """Initialise CUDA (run time only, never at import / backend construction: Terra may fork before that)."""
    try:
        import torch

        if torch.cuda.is_available() and not torch.cuda.is_initialized():
            torch.cuda.init()
    except Exception:  # pragma: no cover
        pass


def _cfg_get(cfg, key, default=None):
    return cfg[key] if key in cfg else default


class QasmSimulatorB200:
    """B200 implementation of the qasm simulator (mirror of QasmSimulatorPy)."""

    NAME = "qasm_simulator"
    BACKEND_VERSION = "2.1.0"
    DESCRIPTION = "A B200 (sm_100a) CUDA simulator for qasm experiments"
    SHOW_FINAL_STATE = False  # qasm_simulator.py:108
    DEFAULT_OPTIONS = {
        "shots": 1024,
        "memory": False,
        "initial_statevector": None,
        "chop_threshold": 1e-15,
        "allow_sample_measuring": True,
        "seed_simulator": None,
        "parameter_binds": None,
    }  # qasm_simulator.py:133-143

    def __init__(self, state_factory=None, max_qubits=None, exact_sampling_max_qubits=40,
                 tile_qubits=None, seed_sync=None, statevector_on_device=False):
        self._state_factory = state_factory or self._device_state
        # multi-process (sharded) runs: every rank executes the same run(qobj); a seed the backend draws itself
        # (qasm_simulator.py:499-502) must then be the same on all ranks -- seed_sync(seed) -> rank 0's seed
        self._seed_sync = seed_sync
        # statevector simulator: leave the final state in HBM and return a statevector.DeviceStatevector handle in
        # data["statevector"] (expectation values / probabilities / sampling without the 2^n-amplitude D2H copy)
        self._statevector_on_device = bool(statevector_on_device)
        self._max_qubits = max_qubits
        self._tile_qubits = tile_qubits
        # up to this many sampled qubits the CDF is accumulated in numpy's cumsum order (bit-exact
        # counts vs the reference; the parallel exact scan costs ~1 % of a circuit even at 33 qubits);
        # above, the blocked parallel scan is used
        self.exact_sampling_max_qubits = exact_sampling_max_qubits
        # persistent option defaults of the owning backend (BackendV1.set_options): the reference resets its
        # per-run options from self.options, not from the class defaults (qasm_simulator.py:292-293)
        self.option_defaults = None
        self._local_random = np.random.RandomState()  # qasm_simulator.py:119
        self._classical_memory = 0
        self._classical_register = 0
        self._state = None
        self._number_of_cmembits = 0
        self._number_of_qubits = 0
        self._shots = 0
        self._memory = False
        self._initial_statevector = None
        self._chop_threshold = 1e-15
        self._qobj_config = None
        self._sample_measure = False
        self.last_stats = {}

    def _device_state(self, n_qubits):
        if self._tile_qubits is not None:
            return DeviceState(n_qubits, tile_qubits=self._tile_qubits)
        return DeviceState(n_qubits)

    # -- configuration (mirror of DEFAULT_CONFIGURATION, qasm_simulator.py:61-102) ---
    def name(self):
        return self.NAME

    def max_qubits(self):
        """The explicit cap if one was given, else what the current GPU holds -- recomputed on every call: before
        CUDA is initialised (backend construction, configuration()) it is the B200 default, at _validate time the
        device's real memory."""
        if self._max_qubits is not None:
            return self._max_qubits
        return max_qubits_for_device()

    def configuration_dict(self):
        gates = [
            {"name": "u1", "parameters": ["lambda"], "qasm_def": "gate u1(lambda) q { U(0,0,lambda) q; }"},
            {"name": "u2", "parameters": ["phi", "lambda"],
             "qasm_def": "gate u2(phi,lambda) q { U(pi/2,phi,lambda) q; }"},
            {"name": "u3", "parameters": ["theta", "phi", "lambda"],
             "qasm_def": "gate u3(theta,phi,lambda) q { U(theta,phi,lambda) q; }"},
            {"name": "rz", "parameters": ["phi"], "qasm_def": "gate rz(phi) q { U(0,0,phi) q; }"},
            {"name": "sx", "parameters": [], "qasm_def": "gate sx(phi) q { U(pi/2,7*pi/2,pi/2) q; }"},
            {"name": "x", "parameters": [], "qasm_def": "gate x q { U(pi,7*pi/2,pi/2) q; }"},
            {"name": "cx", "parameters": [], "qasm_def": "gate cx c,t { CX c,t; }"},
            {"name": "id", "parameters": [], "qasm_def": "gate id a { U(0,0,0) a; }"},
            {"name": "unitary", "parameters": ["matrix"], "qasm_def": "unitary(matrix) q1, q2,..."},
        ]
        return {
            "backend_name": self.NAME,
            "backend_version": self.BACKEND_VERSION,
            "n_qubits": self.max_qubits(),
            "url": "https://github.com/Qiskit/qiskit-terra",
            "simulator": True,
            "local": True,
            "conditional": True,
            "open_pulse": False,
            "memory": True,
            "max_shots": 65536,
            "coupling_map": None,
            "description": self.DESCRIPTION,
            "basis_gates": list(BASIS_GATES),
            "gates": gates,
        }

    # -- options ---------------------------------------------------------------------
    def _set_options(self, qobj_config, backend_options):
        """qasm_simulator.py:289-317."""
        defaults = self.option_defaults if self.option_defaults is not None else self.DEFAULT_OPTIONS
        self._initial_statevector = defaults.get("initial_statevector")
        if self._initial_statevector is not None:
            self._initial_statevector = np.array(self._initial_statevector, dtype=complex)
        self._chop_threshold = defaults.get("chop_threshold")
        backend_options = dict(backend_options or {})
        if "backend_options" in backend_options and backend_options["backend_options"]:
            backend_options = backend_options["backend_options"]
        if "initial_statevector" in backend_options:
            self._initial_statevector = np.array(backend_options["initial_statevector"], dtype=complex)
        elif "initial_statevector" in qobj_config:
            self._initial_statevector = np.array(qobj_config["initial_statevector"], dtype=complex)
        if self._initial_statevector is not None:
            norm = np.linalg.norm(self._initial_statevector)
            if round(norm, 12) != 1:
                raise BasicAerError(
                    "initial statevector is not normalized: " + "norm {} != 1".format(norm)
                )
        if "chop_threshold" in backend_options:
            self._chop_threshold = backend_options["chop_threshold"]
        elif "chop_threshold" in qobj_config:
            self._chop_threshold = qobj_config["chop_threshold"]

    def _validate_initial_statevector(self):
        """qasm_simulator.py:275-287."""
        if self._initial_statevector is None:
            return
        length = len(self._initial_statevector)
        required_dim = 2 ** self._number_of_qubits
        if length != required_dim:
            raise BasicAerError(
                "initial statevector is incorrect length: " + "{} != {}".format(length, required_dim)
            )

    def _validate(self, qobj):
        """qasm_simulator.py:655-675."""
        n_qubits = qobj["config"]["n_qubits"]
        if self._max_qubits is None and self._state_factory == self._device_state:
            _touch_cuda()  # the memory-based cap needs the device's properties
        max_qubits = self.max_qubits()
        if n_qubits > max_qubits:
            raise BasicAerError(
                "Number of qubits {} ".format(n_qubits)
                + "is greater than maximum ({}) ".format(max_qubits)
                + 'for "{}".'.format(self.name())
            )
        for experiment in qobj["experiments"]:
            name = experiment.get("header", {}).get("name")
            if experiment["config"]["memory_slots"] == 0:
                logger.warning('No classical registers in circuit "%s", counts will be empty.', name)
            elif "measure" not in [op["name"] for op in experiment["instructions"]]:
                logger.warning(
                    'No measurements in circuit "%s", classical register will remain all zeros.', name
                )

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

    # -- device operations (each replaces a numpy pass of the reference) ---------------
    def _initialize_statevector(self, global_phase):
        """qasm_simulator.py:319-328 + the global phase multiply at :522, fused."""
        phase = complex(np.exp(1j * global_phase))
        if self._initial_statevector is None:
            self._state.init_basis0(phase)
        else:
            self._state.upload(self._initial_statevector, phase)

    def _flush(self, pending):
        if pending:
            self._state.apply_ops(pending)
            del pending[:]

    def _get_measure_outcome(self, qubit):
        """qasm_simulator.py:163-182; consumes exactly one ``rand()``."""
        probabilities = self._state.prob_qubit(qubit)
        random_number = self._local_random.rand()
        if random_number < probabilities[0]:
            return 0, probabilities[0]
        return 1, probabilities[1]

    def _add_qasm_measure(self, qubit, cmembit, cregbit=None):
        """qasm_simulator.py:227-253."""
        outcome, probability = self._get_measure_outcome(qubit)
        membit = 1 << cmembit
        self._classical_memory = (self._classical_memory & (~membit)) | (outcome << cmembit)
        if cregbit is not None:
            regbit = 1 << cregbit
            self._classical_register = (self._classical_register & (~regbit)) | (outcome << cregbit)
        self._state.collapse(qubit, outcome, probability, is_reset=False)

    def _add_qasm_reset(self, qubit):
        """qasm_simulator.py:255-273."""
        outcome, probability = self._get_measure_outcome(qubit)
        self._state.collapse(qubit, outcome, probability, is_reset=True)

    def _add_sample_measure(self, measure_params, num_samples):
        """qasm_simulator.py:184-225: the marginal, cumsum and searchsorted run on the GPU over the
        uniforms numpy's ``RandomState.choice`` would draw (``random_sample(num_samples)``); the
        bit shuffling into classical memory is vectorised on the host."""
        measured_qubits = sorted({qubit for qubit, _ in measure_params})
        num_measured = len(measured_qubits)
        uniforms = self._local_random.random_sample(num_samples)
        exact = num_measured <= self.exact_sampling_max_qubits
        samples, _total = self._state.sample(measured_qubits, uniforms, exact_order=exact)
        pos = {q: i for i, q in enumerate(measured_qubits)}
        if self._number_of_cmembits <= 62:
            mem = np.full(num_samples, self._classical_memory, dtype=np.int64)
            for qubit, cmembit in measure_params:
                bit = (samples >> pos[qubit]) & 1
                mem = (mem & ~np.int64(1 << cmembit)) | (bit << cmembit)
            return [hex(v) for v in mem.tolist()]
        memory = []
        for sample in samples.tolist():
            classical_memory = self._classical_memory
            for qubit, cmembit in measure_params:
                qubit_outcome = (sample >> pos[qubit]) & 1
                membit = 1 << cmembit
                classical_memory = (classical_memory & (~membit)) | (qubit_outcome << cmembit)
            memory.append(hex(classical_memory))
        return memory

    def _get_statevector(self):
        """qasm_simulator.py:330-334: chop on the device, then one D2H copy owned by the Result."""
        self._state.chop(self._chop_threshold)
        if self._statevector_on_device:
            from .statevector import DeviceStatevector

            return DeviceStatevector(self._state)
        return self._state.download()

    # -- run -------------------------------------------------------------------------
    def run(self, qobj, **backend_options):
        """Run a Qobj dict; returns the result dict (qasm_simulator.py:376-457)."""
        self._set_options(qobj_config=qobj["config"], backend_options=backend_options)
        job_id = str(uuid.uuid4())
        return self._run_job(job_id, qobj)

    def _prepare_qobj(self, qobj):
        return qobj

    def _run_job(self, job_id, qobj):
        qobj = self._prepare_qobj(qobj)
        self._validate(qobj)
        result_list = []
        self._shots = qobj["config"]["shots"]
        self._memory = _cfg_get(qobj["config"], "memory", False)
        self._qobj_config = qobj["config"]
        start = time.time()
        for experiment in qobj["experiments"]:
            result_list.append(self.run_experiment(experiment))
        end = time.time()
        return {
            "backend_name": self.name(),
            "backend_version": self.BACKEND_VERSION,
            "qobj_id": qobj.get("qobj_id"),
            "job_id": job_id,
            "results": result_list,
            "status": "COMPLETED",
            "success": True,
            "time_taken": (end - start),
            "header": qobj.get("header", {}),
        }

    def _passes_condition(self, operation):
        """qasm_simulator.py:527-540."""
        conditional = operation.get("conditional", None)
        if isinstance(conditional, int) and not isinstance(conditional, bool):
            return bool((self._classical_register >> conditional) & 1)
        if conditional is not None:
            mask = int(conditional["mask"], 16)
            if mask > 0:
                value = self._classical_memory & mask
                while (mask & 0x1) == 0:
                    mask >>= 1
                    value >>= 1
                if value != int(conditional["val"], 16):
                    return False
        return True

    def _lower(self, instructions):
        """Pre-lower the gate instructions once per experiment (matrices are shot-independent)."""
        lowered = []
        for operation in instructions:
            name = operation["name"]
            if name == "unitary" or name in lower.SINGLE_QUBIT_GATES or name in ("CX", "cx", "id", "u0",
                                                                                 "barrier"):
                lowered.append(lower.lower_gate(name, operation.get("qubits", []), operation.get("params")))
            else:
                lowered.append(None)
        return lowered

    # -- shot-batched execution of circuits that cannot be sampled --------------------
    MAX_BATCH_STATE_QUBITS = 22   # circuit qubits + log2(shots per batch): 64 MiB of amplitudes
    MIN_BATCH_BITS = 3

    def _shot_batch_bits(self, instructions, lowered):
        """log2 of the shots simulated side by side, or 0 if this experiment runs shot by shot.  Batched: several
        shots, no sampling, a state type with the per-shot kernels (the single-GPU engine), no initial statevector,
        and every measure / reset / bfunc unconditional -- then each shot draws the same number of random numbers and
        the reference's stream (shot after shot, qasm_simulator.py:512-519) can be dealt out in advance."""
        if self._shots < (1 << self.MIN_BATCH_BITS) or self.SHOW_FINAL_STATE or self._initial_statevector is not None:
            return 0
        bits = min(int(np.floor(np.log2(self._shots))), self.MAX_BATCH_STATE_QUBITS - self._number_of_qubits)
        if bits < self.MIN_BATCH_BITS or self._number_of_cmembits > 62:
            return 0
        for operation, low in zip(instructions, lowered):
            name = operation["name"]
            if name in ("measure", "reset", "bfunc"):
                if operation.get("conditional", None) is not None:
                    return 0
            elif low is None and name not in ("id", "u0", "barrier"):
                return 0  # unknown instruction: let the shot-by-shot path raise the reference's error
            elif low is not None and (low.kind == "densek" and operation.get("conditional", None) is not None):
                return 0
        ok = hasattr(self._state, "collapse_batch") and hasattr(self._state, "apply_masked")
        return bits if ok else 0

    def _run_shots_batched(self, instructions, lowered, global_phase, bits):
        """All shots of a non-sampled experiment, 2^bits at a time, as ONE (n + bits)-qubit vector whose top qubits
        number the shots: gates act on every shot through the ordinary fused passes; a measure / reset reads all
        shots' outcome probabilities with one marginal (qubit + shot qubits), decides every shot's outcome with the
        random number the reference would have drawn for it (shot s, k-th draw = element s * D + k of the stream),
        and collapses all shots in one launch; conditional gates run on the shots whose classical bits satisfy the
        condition (qasm_simulator.py:512-614).  Returns the memory list (one hex string per shot, shot order)."""
        from .lower import GateOp

        n = self._number_of_qubits
        batch = 1 << bits
        shot_qubits = list(range(n, n + bits))
        draws = sum(1 for op in instructions if op["name"] in ("measure", "reset"))
        fan = np.array([[1, 0], [1, 0]], dtype=complex)  # |0> -> |0> + |1>: one copy of |0..0> per shot
        phase = complex(np.exp(1j * global_phase))
        memory = []
        done = 0
        self._state = self._state_factory(n + bits)
        while done < self._shots:
            live = min(batch, self._shots - done)
            rand = self._local_random.rand(live * draws).reshape(live, draws) if draws else None
            cmem = np.zeros(batch, dtype=np.int64)
            creg = np.zeros(batch, dtype=np.int64)
            self._state.init_basis0(phase)
            pending = [GateOp(lower._lib.GATE_DENSE1, [q], fan) for q in shot_qubits]
            draw = 0
            for operation, low in zip(instructions, lowered):
                name = operation["name"]
                conditional = operation.get("conditional", None)
                if low is not None:
                    if conditional is None:
                        pending.append(low)
                    else:
                        self._flush(pending)
                        mask = self._condition_mask(conditional, cmem, creg)
                        if mask.any():
                            self._state.apply_masked(bits, low, mask)
                elif name in ("id", "u0", "barrier"):
                    pass
                elif name in ("measure", "reset"):
                    self._flush(pending)
                    qubit = operation["qubits"][0]
                    probs = self._state.marginal([qubit] + shot_qubits).cpu().numpy().reshape(batch, 2)
                    r = np.ones(batch)
                    r[:live] = rand[:, draw]
                    draw += 1
                    outcome = np.where(r < probs[:, 0], 0, 1)           # _get_measure_outcome :178-182
                    outcome = np.where(probs[np.arange(batch), outcome] > 0.0, outcome, 1 - outcome)  # unused slots
                    chosen = probs[np.arange(batch), outcome]
                    self._state.collapse_batch(bits, qubit, outcome, 1.0 / np.sqrt(chosen), is_reset=(name == "reset"))
                    if name == "measure":
                        cmembit = operation["memory"][0]
                        cmem = (cmem & ~np.int64(1 << cmembit)) | (outcome.astype(np.int64) << cmembit)
                        if "register" in operation:
                            cregbit = operation["register"][0]
                            creg = (creg & ~np.int64(1 << cregbit)) | (outcome.astype(np.int64) << cregbit)
                elif name == "bfunc":
                    mask = int(operation["mask"], 16)
                    val = int(operation["val"], 16)
                    compared = (creg & mask) - val
                    relation = operation["relation"]
                    if relation == "==":
                        out = compared == 0
                    elif relation == "!=":
                        out = compared != 0
                    elif relation == "<":
                        out = compared < 0
                    elif relation == "<=":
                        out = compared <= 0
                    elif relation == ">":
                        out = compared > 0
                    elif relation == ">=":
                        out = compared >= 0
                    else:
                        raise BasicAerError("Invalid boolean function relation.")
                    cregbit = operation["register"]
                    creg = (creg & ~np.int64(1 << cregbit)) | (out.astype(np.int64) << cregbit)
                    cmembit = operation.get("memory", None)
                    if cmembit is not None:
                        cmem = (cmem & ~np.int64(1 << cmembit)) | (out.astype(np.int64) << cmembit)
            self._flush(pending)
            memory.extend(hex(int(v)) for v in cmem[:live])
            done += live
        return memory

    @staticmethod
    def _condition_mask(conditional, cmem, creg):
        """_passes_condition (qasm_simulator.py:527-540) for every shot of a batch."""
        if isinstance(conditional, int) and not isinstance(conditional, bool):
            return ((creg >> conditional) & 1).astype(bool)
        mask = int(conditional["mask"], 16)
        if mask <= 0:
            return np.ones(cmem.size, dtype=bool)
        value = cmem & mask
        while (mask & 0x1) == 0:
            mask >>= 1
            value = value >> 1
        return value == int(conditional["val"], 16)

    def run_experiment(self, experiment):
        """qasm_simulator.py:459-653."""
        start = time.time()
        exp_config = experiment.get("config", {})
        header = experiment.get("header", {})
        instructions = experiment["instructions"]
        self._number_of_qubits = exp_config["n_qubits"]
        self._number_of_cmembits = exp_config["memory_slots"]
        self._classical_memory = 0
        self._classical_register = 0
        self._sample_measure = False
        global_phase = header.get("global_phase", 0.0)
        self._validate_initial_statevector()
        if "seed_simulator" in exp_config:
            seed_simulator = exp_config["seed_simulator"]
        elif "seed_simulator" in self._qobj_config:
            seed_simulator = self._qobj_config["seed_simulator"]
        else:
            seed_simulator = np.random.randint(2147483647, dtype="int32")
            if self._seed_sync is not None:
                seed_simulator = self._seed_sync(int(seed_simulator))
        self._local_random.seed(seed=seed_simulator)
        self._validate_measure_sampling(experiment)

        try:
            self._state = self._state_factory(self._number_of_qubits)
        except MemoryError as exc:
            raise BasicAerError("not enough device memory for {} qubits: {}".format(self._number_of_qubits, exc))
        except RuntimeError as exc:  # torch.cuda.OutOfMemoryError is a RuntimeError
            if "out of memory" not in str(exc).lower():
                raise
            raise BasicAerError("not enough device memory for {} qubits: {}".format(self._number_of_qubits, exc))
        lowered = self._lower(instructions)

        memory = []
        if self._sample_measure:
            shots = 1
            measure_sample_ops = []
        else:
            shots = self._shots

        batch_bits = 0 if self._sample_measure else self._shot_batch_bits(instructions, lowered)
        if batch_bits:
            memory = self._run_shots_batched(instructions, lowered, global_phase, batch_bits)
            shots = 0  # the per-shot loop below is skipped

        # The gate-only, unconditional prefix of the circuit is shot-independent and draws no random
        # numbers: simulate it once and restore it for every further shot.
        prefix_len = 0
        for operation, low in zip(instructions, lowered):
            if operation.get("conditional", None) is not None or operation["name"] in (
                    "measure", "reset", "bfunc") or (low is None and operation["name"] not in
                                                      ("id", "u0", "barrier")):
                break
            prefix_len += 1
        snapshot = None

        for shot in range(shots):
            self._classical_memory = 0
            self._classical_register = 0
            pending = []
            if snapshot is not None:
                self._state.restore(snapshot)
                first = prefix_len
            else:
                self._initialize_statevector(global_phase)
                first = 0
                if shots > 1 and prefix_len > 0 and self._number_of_qubits <= 31:
                    pending.extend(op for op in lowered[:prefix_len] if op is not None)
                    self._flush(pending)
                    snapshot = self._state.snapshot()
                    first = prefix_len
            for operation, low in zip(instructions[first:], lowered[first:]):
                if not self._passes_condition(operation):
                    continue
                name = operation["name"]
                if low is not None:
                    pending.append(low)
                elif name in ("id", "u0", "barrier"):
                    pass
                elif name == "reset":
                    self._flush(pending)
                    self._add_qasm_reset(operation["qubits"][0])
                elif name == "measure":
                    qubit = operation["qubits"][0]
                    cmembit = operation["memory"][0]
                    cregbit = operation["register"][0] if "register" in operation else None
                    if self._sample_measure:
                        measure_sample_ops.append((qubit, cmembit))
                    else:
                        self._flush(pending)
                        self._add_qasm_measure(qubit, cmembit, cregbit)
                elif name == "bfunc":
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
                        raise BasicAerError("Invalid boolean function relation.")
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
                    backend = self.name()
                    err_msg = '{0} encountered unrecognized operation "{1}"'
                    raise BasicAerError(err_msg.format(backend, name))
            self._flush(pending)

            if self._number_of_cmembits > 0:
                if self._sample_measure:
                    memory = self._add_sample_measure(measure_sample_ops, self._shots)
                else:
                    memory.append(hex(self._classical_memory))

        if batch_bits and self._number_of_cmembits == 0:
            memory = []
        data = {"counts": dict(Counter(memory))}
        if self._memory:
            data["memory"] = memory
        if self.SHOW_FINAL_STATE:
            data["statevector"] = self._get_statevector()
            if not data["counts"]:
                data.pop("counts")
            if "memory" in data and not data["memory"]:
                data.pop("memory")
        self.last_stats = {
            "passes": getattr(self._state, "passes_launched", None),
            "gates": getattr(self._state, "gates_applied", None),
            "shot_batch_bits": batch_bits,
        }
        self._state = None
        end = time.time()
        return {
            "name": header.get("name"),
            "seed_simulator": seed_simulator,
            "shots": self._shots,
            "data": data,
            "status": "DONE",
            "success": True,
            "time_taken": (end - start),
            "header": header,
        }


class StatevectorSimulatorB200(QasmSimulatorB200):
    """Mirror of StatevectorSimulatorPy (statevector_simulator.py:36-122)."""

    NAME = "statevector_simulator"
    BACKEND_VERSION = "1.1.0"
    DESCRIPTION = "A B200 (sm_100a) CUDA statevector simulator for qobj files"
    SHOW_FINAL_STATE = True  # statevector_simulator.py:85

    def _prepare_qobj(self, qobj):
        """statevector_simulator.py:111-122: only one shot is supported."""
        if qobj["config"].get("shots", 1) != 1:
            logger.info('"%s" only supports 1 shot. Setting shots=1.', self.name())
            qobj["config"]["shots"] = 1
        for experiment in qobj["experiments"]:
            if experiment.get("config", {}).get("shots", 1) != 1:
                experiment["config"]["shots"] = 1
        return qobj


class UnitarySimulatorB200:
    """Mirror of UnitarySimulatorPy (unitary_simulator.py:57-401).

    The unitary is held as a 2n-"qubit" device vector with flat index row*2^n + col; a gate on qubit q
    acts on bit n+q (the row bits), exactly the einsum of unitary_simulator.py:127-141."""

    NAME = "unitary_simulator"
    BACKEND_VERSION = "1.1.0"
    DESCRIPTION = "A B200 (sm_100a) CUDA unitary simulator for qobj files"
    DEFAULT_OPTIONS = {"shots": 1, "initial_unitary": None, "chop_threshold": 1e-15,
                       "parameter_binds": None}  # unitary_simulator.py:123-125

    def __init__(self, state_factory=None, max_qubits=None):
        self._state_factory = state_factory or DeviceState
        self._max_qubits = max_qubits
        self.option_defaults = None  # see QasmSimulatorB200.option_defaults (unitary_simulator.py:160-161)
        self._initial_unitary = None
        self._chop_threshold = 1e-15
        self._global_phase = 0
        self._number_of_qubits = 0

    def name(self):
        return self.NAME

    def max_qubits(self):
        if self._max_qubits is not None:
            return self._max_qubits
        return max_qubits_for_device() // 2

    def configuration_dict(self):
        cfg = QasmSimulatorB200(max_qubits=self.max_qubits()).configuration_dict()
        cfg.update({"backend_name": self.NAME, "backend_version": self.BACKEND_VERSION,
                    "description": self.DESCRIPTION, "n_qubits": self.max_qubits(),
                    "conditional": False, "memory": False})
        return cfg

    def _set_options(self, qobj_config, backend_options):
        """unitary_simulator.py:157-188."""
        defaults = self.option_defaults if self.option_defaults is not None else self.DEFAULT_OPTIONS
        self._initial_unitary = defaults.get("initial_unitary")
        if self._initial_unitary is not None:
            self._initial_unitary = np.array(self._initial_unitary, dtype=complex)
        self._chop_threshold = defaults.get("chop_threshold")
        backend_options = dict(backend_options or {})
        if "backend_options" in backend_options:
            backend_options = backend_options["backend_options"]
        if "initial_unitary" in backend_options:
            self._initial_unitary = np.array(backend_options["initial_unitary"], dtype=complex)
        elif "initial_unitary" in qobj_config:
            self._initial_unitary = np.array(qobj_config["initial_unitary"], dtype=complex)
        if self._initial_unitary is not None:
            shape = np.shape(self._initial_unitary)
            if len(shape) != 2 or shape[0] != shape[1]:
                raise BasicAerError("initial unitary is not a square matrix")
            iden = np.eye(len(self._initial_unitary))
            u_dagger_u = np.dot(self._initial_unitary.T.conj(), self._initial_unitary)
            norm = np.linalg.norm(u_dagger_u - iden)
            if round(norm, 10) != 0:
                raise BasicAerError("initial unitary is not unitary")
        if "chop_threshold" in backend_options:
            self._chop_threshold = backend_options["chop_threshold"]
        elif "chop_threshold" in qobj_config:
            self._chop_threshold = qobj_config["chop_threshold"]

    def _validate(self, qobj):
        """unitary_simulator.py:368-401."""
        n_qubits = qobj["config"]["n_qubits"]
        max_qubits = self.max_qubits()
        if n_qubits > max_qubits:
            raise BasicAerError(
                "Number of qubits {} ".format(n_qubits)
                + "is greater than maximum ({}) ".format(max_qubits)
                + 'for "{}".'.format(self.name())
            )
        if "shots" in qobj["config"] and qobj["config"]["shots"] != 1:
            logger.info('"%s" only supports 1 shot. Setting shots=1.', self.name())
            qobj["config"]["shots"] = 1
        for experiment in qobj["experiments"]:
            name = experiment.get("header", {}).get("name")
            if experiment.get("config", {}).get("shots", 1) != 1:
                experiment["config"]["shots"] = 1
            for operation in experiment["instructions"]:
                if operation["name"] in ["measure", "reset"]:
                    raise BasicAerError(
                        'Unsupported "%s" instruction "%s" ' + 'in circuit "%s" ',
                        self.name(),
                        operation["name"],
                        name,
                    )

    def run(self, qobj, **backend_options):
        """unitary_simulator.py:209-292."""
        self._set_options(qobj_config=qobj["config"], backend_options=backend_options)
        job_id = str(uuid.uuid4())
        self._validate(qobj)
        result_list = []
        start = time.time()
        for experiment in qobj["experiments"]:
            result_list.append(self.run_experiment(experiment))
        end = time.time()
        return {
            "backend_name": self.name(),
            "backend_version": self.BACKEND_VERSION,
            "qobj_id": qobj.get("qobj_id"),
            "job_id": job_id,
            "results": result_list,
            "status": "COMPLETED",
            "success": True,
            "time_taken": (end - start),
            "header": qobj.get("header", {}),
        }

    def run_experiment(self, experiment):
        """unitary_simulator.py:294-366."""
        start = time.time()
        header = experiment.get("header", {})
        n = header.get("n_qubits", experiment["config"]["n_qubits"])
        self._number_of_qubits = n
        self._global_phase = header.get("global_phase", 0.0)
        if self._initial_unitary is not None:  # _validate_initial_unitary :143-155
            shape = np.shape(self._initial_unitary)
            required_shape = (2 ** n, 2 ** n)
            if shape != required_shape:
                raise BasicAerError(
                    "initial unitary is incorrect shape: " + "{} != 2 ** {}".format(shape, required_shape)
                )
        state = self._state_factory(2 * n)
        # the reference multiplies by the global phase at read-out (:204-205); a scalar commutes with
        # every gate, so it is folded into the initialisation here
        phase = complex(np.exp(1j * float(self._global_phase))) if self._global_phase else 1.0 + 0.0j
        if self._initial_unitary is None:
            state.init_identity(n, phase)
        else:
            state.upload(self._initial_unitary.reshape(-1), phase)
        pending = []
        for operation in experiment["instructions"]:
            name = operation["name"]
            if name == "unitary" or name in lower.SINGLE_QUBIT_GATES or name in ("CX", "cx"):
                op = lower.lower_gate(name, operation["qubits"], operation.get("params"))
                pending.append(lower.shift_op(op, n))
            elif name in ("id", "u0", "barrier"):
                pass
            else:
                backend = self.name()
                err_msg = '{0} encountered unrecognized operation "{1}"'
                raise BasicAerError(err_msg.format(backend, name))
        state.apply_ops(pending)
        state.chop(self._chop_threshold)
        unitary = state.download().reshape(2 ** n, 2 ** n)
        end = time.time()
        return {
            "name": header.get("name"),
            "shots": 1,
            "data": {"unitary": unitary},
            "status": "DONE",
            "success": True,
            "time_taken": (end - start),
            "header": header,
        }


def get_simulator(name, **kwargs):
    """Backend lookup including the deprecated aliases (basicaerprovider.py:79-89)."""
    aliases = {
        "qasm_simulator_py": "qasm_simulator",
        "statevector_simulator_py": "statevector_simulator",
        "unitary_simulator_py": "unitary_simulator",
        "local_qasm_simulator_py": "qasm_simulator",
        "local_statevector_simulator_py": "statevector_simulator",
        "local_unitary_simulator_py": "unitary_simulator",
        "local_unitary_simulator": "unitary_simulator",
    }
    name = aliases.get(name, name)
    classes = {"qasm_simulator": QasmSimulatorB200, "statevector_simulator": StatevectorSimulatorB200,
               "unitary_simulator": UnitarySimulatorB200}
    if name not in classes:
        raise LookupError("The '%s' backend is not installed in your system." % name)
    return classes[name](**kwargs)
