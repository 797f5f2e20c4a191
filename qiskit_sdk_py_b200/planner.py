"""Fusion planner: a stream of 1-/2-qubit gate ops -> cache-blocked passes.

The reference applies one gate per full pass over the vector (the instruction loop at
qiskit/providers/basicaer/qasm_simulator.py:526-559 calls _add_unitary -> np.einsum per op).  Here a
*pass* stages tiles of 2^b amplitudes in shared memory, so every gate whose qubits lie inside the
tile's b "tile qubits" can be applied in the same HBM round trip.  The planner picks, greedily and in
program order, the set of gates for each pass:

  * a gate may join the pass if its qubits fit in the remaining tile budget and none of its qubits
    has been touched by an earlier gate that was left for a later pass (so the relative order of
    gates that share a qubit never changes -- only gates on disjoint qubits are commuted);
  * the low qubits 0..3 are always tile qubits, so every global-memory access of the pass is part of a
    256-byte contiguous run (SURVEY 2.5 K4 / "Hard parts: low-qubit gates"; measured copy rates of the
    pass kernel on B200: 128-byte runs 4.06 TB/s, 256-byte runs 5.03 TB/s, fully contiguous 5.27 TB/s).

Pure host logic, no device access: unit-tested on CPU.
"""
from . import _lib

DEFAULT_TILE_QUBITS = 11
LOW_QUBITS = 4  # 16 amplitudes = 256 B contiguous runs (measured: 128 B runs 4.06 TB/s, 256 B 5.03, contiguous 5.27)


class Pass:
    __slots__ = ("tile_qubits", "ops")

    def __init__(self, tile_qubits, ops):
        self.tile_qubits = tuple(tile_qubits)
        self.ops = list(ops)

    def __repr__(self):
        return "Pass(tile=%r, %d ops)" % (self.tile_qubits, len(self.ops))


def _fill_tile(n_qubits, used, tile_size):
    """used gate qubits + the lowest other qubits, up to tile_size."""
    tile = set(used)
    q = 0
    while len(tile) < tile_size and q < n_qubits:
        tile.add(q)
        q += 1
    return sorted(tile)


def plan_passes(ops, n_qubits, tile_qubits=DEFAULT_TILE_QUBITS, max_gates=_lib.MAX_PASS_GATES,
                max_payload=_lib.MAX_PASS_MATRIX_DOUBLES, window=4096):
    """Group ``ops`` (GateOps with <= 2 qubits each) into passes.  Returns a list of Pass."""
    tile_size = min(n_qubits, tile_qubits, _lib.MAX_TILE_QUBITS)
    low = set(range(min(LOW_QUBITS, n_qubits)))
    for op in ops:
        if len(op.qubits) > 2 or op.kind == "densek":
            raise ValueError("plan_passes only fuses 1- and 2-qubit gates")
        if len(set(op.qubits) | low) > tile_size:
            # cannot keep the low qubits in the tile for this gate (tiny tile sizes only)
            low = set()
    remaining = list(ops)
    passes = []
    while remaining:
        used = set(low)
        blocked = set()
        chosen = []
        payload = 0
        keep = []
        scanned = 0
        for idx, op in enumerate(remaining):
            qs = set(op.qubits)
            take = False
            if scanned < window and not (qs & blocked) and len(chosen) < max_gates:
                if len(used | qs) <= tile_size and payload + op.payload_doubles <= max_payload:
                    take = True
            scanned += 1
            if take:
                used |= qs
                chosen.append(op)
                payload += op.payload_doubles
            else:
                blocked |= qs
                keep.append(op)
                if len(blocked) >= n_qubits or scanned >= window:
                    keep.extend(remaining[idx + 1:])
                    break
        if not chosen:  # cannot happen: the first op always fits
            raise RuntimeError("planner made no progress")
        passes.append(Pass(_fill_tile(n_qubits, used, tile_size), chosen))
        remaining = keep
    return passes
