"""Benchmark / parity circuit generators, emitted directly as Qobj dicts in the simulator basis.

The reference builds these circuits with ``QuantumCircuit`` + ``transpile`` and serialises them
with ``assemble`` (qiskit/compiler/assembler.py:43); the wire format that reaches the simulator
is ``QasmQobj.to_dict()`` (qiskit/qobj/qasm_qobj.py:30-102,203-330,538-650).  These generators
write that format directly, so they work without qiskit (the GPU box has none) and without
``retworkx`` (``transpile`` cannot run in the build container).  ``tests/golden/make_golden.py``
checks them against circuits built with the reference's own classes.

Gate decompositions used (all from the reference's gate definitions):
  h        = u2(0, pi)                                  standard_gates/h.py:54-68
  cp(l)a,b = u1(l/2) a; cx a,b; u1(-l/2) b; cx a,b; u1(l/2) b   standard_gates/p.py:168-186
  swap a,b = cx a,b; cx b,a; cx a,b                     standard_gates/swap.py:57-75
"""
import math

import numpy as np


def _header(name, n_qubits, memory_slots, global_phase=0.0, creg_sizes=None):
    creg_sizes = creg_sizes if creg_sizes is not None else ([["c", memory_slots]] if memory_slots else [])
    clbit_labels = []
    for reg, size in creg_sizes:
        clbit_labels += [[reg, i] for i in range(size)]
    return {
        "qubit_labels": [["q", i] for i in range(n_qubits)],
        "n_qubits": n_qubits,
        "qreg_sizes": [["q", n_qubits]],
        "clbit_labels": clbit_labels,
        "memory_slots": memory_slots,
        "creg_sizes": creg_sizes,
        "name": name,
        "global_phase": global_phase,
        "metadata": {},
    }


def make_experiment(name, n_qubits, instructions, memory_slots=0, global_phase=0.0, creg_sizes=None,
                    config=None):
    """One Qobj experiment dict (qiskit/assembler/assemble_circuits.py:36-163 layout)."""
    cfg = {"n_qubits": n_qubits, "memory_slots": memory_slots}
    if config:
        cfg.update(config)
    return {
        "config": cfg,
        "header": _header(name, n_qubits, memory_slots, global_phase, creg_sizes),
        "instructions": list(instructions),
    }


def make_qobj(experiments, shots=1024, seed_simulator=None, memory=False, qobj_id="b200-qobj", **config):
    """A QASM Qobj dict (qiskit/assembler/assemble_circuits.py:320-345, run_config.py:51-61:
    ``seed_simulator`` is absent, not None, when the caller gave none)."""
    if isinstance(experiments, dict):
        experiments = [experiments]
    cfg = {
        "shots": shots,
        "memory": memory,
        "parameter_binds": [],
        "meas_level": 2,
        "init_qubits": True,
        "memory_slots": max(e["config"]["memory_slots"] for e in experiments),
        "n_qubits": max(e["config"]["n_qubits"] for e in experiments),
    }
    if seed_simulator is not None:
        cfg["seed_simulator"] = seed_simulator
    cfg.update(config)
    return {
        "qobj_id": qobj_id,
        "header": {},
        "config": cfg,
        "schema_version": "1.3.0",
        "type": "QASM",
        "experiments": list(experiments),
    }


# -- instruction helpers ----------------------------------------------------------------
def u1(lam, q):
    return {"name": "u1", "params": [float(lam)], "qubits": [q]}


def u2(phi, lam, q):
    return {"name": "u2", "params": [float(phi), float(lam)], "qubits": [q]}


def u3(theta, phi, lam, q):
    return {"name": "u3", "params": [float(theta), float(phi), float(lam)], "qubits": [q]}


def cx(c, t):
    return {"name": "cx", "qubits": [c, t]}


def unitary(matrix, qubits):
    return {"name": "unitary", "params": [np.asarray(matrix, dtype=complex)], "qubits": list(qubits)}


def measure(q, c):
    return {"name": "measure", "qubits": [q], "memory": [c]}


def measure_all(n):
    return [measure(q, q) for q in range(n)]


# -- circuits ---------------------------------------------------------------------------
def bell(shots=1024, seed_simulator=42):
    """BASELINE config 1: u2(0,pi) q0; cx q0,q1; measure (README.md:32-41 in basis gates)."""
    ins = [u2(0.0, math.pi, 0), cx(0, 1), measure(0, 0), measure(1, 1)]
    return make_qobj(make_experiment("bell", 2, ins, memory_slots=2), shots=shots,
                     seed_simulator=seed_simulator)


def qft_instructions(n, do_swaps=True):
    """QFT in basis gates; gate order follows QFT._build
    (qiskit/circuit/library/basis_change/qft.py:252-272)."""
    ins = []
    for j in reversed(range(n)):
        ins.append(u2(0.0, math.pi, j))
        for k in reversed(range(0, j)):
            lam = math.pi / (2 ** (j - k))
            # cp(lam, j, k): control j, target k
            ins += [u1(lam / 2, j), cx(j, k), u1(-lam / 2, k), cx(j, k), u1(lam / 2, k)]
    if do_swaps:
        for i in range(n // 2):
            a, b = i, n - i - 1
            ins += [cx(a, b), cx(b, a), cx(a, b)]
    return ins


def product_state_layer(n, seed):
    """A seeded u3 layer so QFT acts on a non-trivial product state."""
    rng = np.random.RandomState(seed)
    return [u3(*(rng.uniform(0, 2 * math.pi, 3)), q) for q in range(n)]


def qft(n, state_seed=None, shots=1, measure_final=False, seed_simulator=None):
    """BASELINE config 2 (n=20): 20 u2 + 190*(3 u1 + 2 cx) + 30 cx = 1000 ops."""
    ins = (product_state_layer(n, state_seed) if state_seed is not None else []) + qft_instructions(n)
    slots = 0
    if measure_final:
        ins += measure_all(n)
        slots = n
    return make_qobj(make_experiment("qft_%d" % n, n, ins, memory_slots=slots), shots=shots,
                     seed_simulator=seed_simulator)


def quantum_volume_instructions(n, depth=None, seed=1234):
    """SU(4) layers exactly as QuantumVolume(n, depth, seed) generates them
    (qiskit/circuit/library/quantum_volume.py:76-113): ``default_rng(seed)`` draws the
    per-gate unitary seeds first, then one permutation per layer; each SU(4) is
    ``scipy.stats.unitary_group.rvs(4, random_state=default_rng(seed_u))``
    (qiskit/quantum_info/operators/random.py:44-57)."""
    from scipy import stats

    rng = np.random.default_rng(seed)
    depth = depth or n
    width = n // 2
    unitary_seeds = rng.integers(low=1, high=1000, size=[depth, width])
    perm_0 = list(range(n))
    ins = []
    cache = {}
    for d in range(depth):
        perm = rng.permutation(perm_0)
        for w in range(width):
            seed_u = int(unitary_seeds[d][w])
            if seed_u not in cache:
                cache[seed_u] = stats.unitary_group.rvs(4, random_state=np.random.default_rng(seed_u))
            ins.append(unitary(cache[seed_u], [int(perm[2 * w]), int(perm[2 * w + 1])]))
    return ins


def quantum_volume(n, depth=None, seed=1234, shots=8192, seed_simulator=42, measure_final=True,
                   memory=False):
    """BASELINE configs 3/4: QuantumVolume(n, depth, seed) + measure_all."""
    ins = quantum_volume_instructions(n, depth, seed)
    slots = 0
    if measure_final:
        ins += measure_all(n)
        slots = n
    name = "quantum_volume_[%d,%d,%d]" % (n, depth or n, seed)
    return make_qobj(make_experiment(name, n, ins, memory_slots=slots), shots=shots,
                     seed_simulator=seed_simulator, memory=memory)


def random_basis_instructions(n, depth, seed, two_qubit_prob=0.4, with_unitary=True):
    """Random layers drawn from the simulator basis u1,u2,u3,rz,sx,x,cx,id,unitary
    (the basis in qiskit/providers/basicaer/qasm_simulator.py:74), in the spirit of
    random_circuit (qiskit/circuit/random/utils.py:55-175)."""
    rng = np.random.RandomState(seed)
    ins = []
    for _ in range(depth):
        perm = rng.permutation(n)
        i = 0
        while i < n:
            if i + 1 < n and rng.rand() < two_qubit_prob:
                a, b = int(perm[i]), int(perm[i + 1])
                if with_unitary and rng.rand() < 0.5:
                    z = rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))
                    q, r = np.linalg.qr(z)
                    ins.append(unitary(q * (np.diag(r) / np.abs(np.diag(r))), [a, b]))
                else:
                    ins.append(cx(a, b))
                i += 2
            else:
                q = int(perm[i])
                kind = rng.randint(7)
                ang = rng.uniform(0, 2 * math.pi, 3)
                if kind == 0:
                    ins.append(u1(ang[0], q))
                elif kind == 1:
                    ins.append(u2(ang[0], ang[1], q))
                elif kind == 2:
                    ins.append(u3(ang[0], ang[1], ang[2], q))
                elif kind == 3:
                    ins.append({"name": "rz", "params": [float(ang[0])], "qubits": [q]})
                elif kind == 4:
                    ins.append({"name": "sx", "qubits": [q]})
                elif kind == 5:
                    ins.append({"name": "x", "qubits": [q]})
                else:
                    ins.append({"name": "id", "qubits": [q]})
                i += 1
    return ins


def random_circuit(n, depth, seed, shots=1024, seed_simulator=7, measure_final=True, memory=False):
    ins = random_basis_instructions(n, depth, seed)
    slots = 0
    if measure_final:
        ins += measure_all(n)
        slots = n
    return make_qobj(make_experiment("random_%d_%d_%d" % (n, depth, seed), n, ins, memory_slots=slots),
                     shots=shots, seed_simulator=seed_simulator, memory=memory)


def count_gate_passes(instructions):
    """Number of reference gate passes (one np.einsum each, qasm_simulator.py:159) in a stream."""
    return sum(1 for op in instructions if op["name"] in
               ("unitary", "U", "u1", "u2", "u3", "rz", "sx", "x", "cx", "CX"))
