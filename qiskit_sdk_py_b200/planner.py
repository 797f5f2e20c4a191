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
import numpy as np

from . import _lib
from .lower import GateOp, matrix_op

DEFAULT_TILE_QUBITS = 11
import os as _os

# 16 amplitudes = 256 B contiguous runs (measured: 128 B runs 4.06 TB/s, 256 B 5.03, contiguous 5.27).
# QSV_LOW_QUBITS overrides it for planner experiments (fewer pinned qubits = fewer passes, smaller TMA pieces).
try:
    LOW_QUBITS = min(4, max(0, int(_os.environ.get("QSV_LOW_QUBITS", "4"))))
except ValueError:
    LOW_QUBITS = 4


class Pass:
    """One launch of the pass kernel: ``tile_qubits`` and the qubits of ``ops`` are PHYSICAL bit positions
    at the start of the pass; ``dest`` (or None) says where each tile qubit's content lands on
    write-back (``dest[i]`` for ``tile_qubits[i]``)."""

    __slots__ = ("tile_qubits", "ops", "dest", "tail")

    def __init__(self, tile_qubits, ops, dest=None, tail=None):
        self.tile_qubits = tuple(tile_qubits)
        self.ops = list(ops)
        self.dest = None if dest is None else tuple(dest)
        self.tail = list(tail) if tail else []  # diagonal ops on any qubits, applied after ``ops`` in the same pass

    def __repr__(self):
        return "Pass(tile=%r, %d ops%s%s)" % (self.tile_qubits, len(self.ops), ", remap" if self.dest else "",
                                              ", %d tail" % len(self.tail) if self.tail else "")


_CX = np.array([[1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0], [0, 1, 0, 0]], dtype=complex)  # basicaertools.py:70-72


def _dense(op):
    """The op as a dense matrix in its own qubit order (qubits[0] = LSB of the index)."""
    if op.kind in (_lib.GATE_DIAG1, _lib.GATE_DIAG2):
        return np.diag(np.asarray(op.matrix, dtype=complex))
    if op.kind == _lib.GATE_CX:
        return _CX
    return np.asarray(op.matrix, dtype=complex)


_SWAP_BITS = np.array([0, 2, 1, 3])


def _embed(mat, qubits, support):
    """``mat`` on ``qubits`` as a matrix on ``support`` (1 or 2 qubits, support[0] = LSB).  Written with slice
    assignments: this runs once per merged gate on the host, np.kron would dominate small circuits."""
    if qubits == support:
        return mat
    if len(qubits) == 2:  # same pair, roles exchanged: swap the two index bits
        return mat[_SWAP_BITS][:, _SWAP_BITS]
    out = np.zeros((4, 4), dtype=complex)
    if support[0] == qubits[0]:  # acts on index bit 0: I (x) M
        out[0:2, 0:2] = mat
        out[2:4, 2:4] = mat
    else:                        # acts on index bit 1: M (x) I
        out[0::2, 0::2] = mat
        out[1::2, 1::2] = mat
    return out


def merge_ops(ops):
    """Peephole fusion on the host: a gate is multiplied into the latest earlier gate when together they act
    on at most two qubits and no gate in between touches those qubits.  A controlled-phase written in the
    simulator basis (u1 a; cx a,b; u1 b; cx a,b; u1 b -- five passes over the tile in shared memory, and five
    einsum passes over the whole vector in the reference, qasm_simulator.py:145-161) becomes ONE diagonal
    2-qubit op; runs of 1-qubit gates collapse; a 1-qubit gate next to a 2-qubit gate on its qubit is free.
    The product is re-classified (diagonal / dense), so exact zeros keep diagonal gates on the cheap path.
    Ops that merge with nothing are passed through untouched (same objects)."""
    out = []    # [support, matrix or None (= untouched original), original op]
    last = {}   # qubit -> index in ``out`` of the latest op touching it
    for op in ops:
        qs = tuple(op.qubits)
        if op.kind == "densek" or len(qs) > 2:
            for q in qs:
                last[q] = len(out)
            out.append([qs, None, op])
            continue
        k = max((last.get(q, -1) for q in qs), default=-1)
        target = out[k] if k >= 0 else None
        if target is not None and target[2].kind != "densek" and len(target[0]) <= 2:
            # k is the latest op on ANY qubit of the incoming gate, so nothing after position k touches the
            # gate's qubits: it commutes back to position k, where it multiplies the target
            union = target[0] + tuple(q for q in qs if q not in target[0])
            if len(union) <= 2:
                prev = target[1] if target[1] is not None else _dense(target[2])
                mat = _embed(_dense(op), qs, union) @ _embed(prev, target[0], union)
                target[0], target[1] = union, mat
                for q in qs:
                    last[q] = k
                continue
        for q in qs:
            last[q] = len(out)
        out.append([qs, None, op])
    merged = []
    for support, mat, op in out:
        merged.append(op if mat is None else matrix_op(mat, support))
    return merged


_DIAG_KINDS = (_lib.GATE_DIAG1, _lib.GATE_DIAG2)
_RUN_BITS = 8  # tile bits one diagonal lookup table of the pass kernel can span (tile_pass.cu kRunBits)


def schedule_pass_ops(ops):
    """Reorder the gates of ONE pass (all inside the tile) for the kernel's lowering, which fuses a dense 2-qubit
    gate with the gates that FOLLOW it inside (at most) three qubits into one 3-qubit block and folds CONSECUTIVE
    diagonal gates into one table sweep.  List scheduling over the dependency DAG of the pass -- two gates depend on
    each other iff they share a qubit and are not both diagonal (diagonal gates commute with each other):

      * while a non-diagonal gate is ready, take the first one (program order) and let it pull in, one after the
        other, every ready gate the block can absorb (a gate inside the block's qubits, or a 2-qubit gate that
        extends a 2-qubit block to three qubits);
      * when only diagonal gates are ready (the next dense gates wait for them), take ALL of them at once.

    A QFT pass (Hadamard-merged dense gates separated by controlled-phase fans) goes from 6 dense ops + 5 diagonal
    runs to 3 blocks + 2 runs; streams without diagonal gates keep the blocks that program order gives, or better.
    The returned order is a topological order of that DAG, hence the same unitary."""
    n = len(ops)
    if n < 3:
        return list(ops)
    qsets = [frozenset(op.qubits) for op in ops]
    diag = [op.kind in _DIAG_KINDS for op in ops]
    succs = [[] for _ in range(n)]
    indeg = [0] * n
    for j in range(n):
        for i in range(j):
            if (qsets[i] & qsets[j]) and not (diag[i] and diag[j]):
                succs[i].append(j)
                indeg[j] += 1
    ready = set(i for i in range(n) if indeg[i] == 0)
    out = []

    def emit(i):
        ready.discard(i)
        out.append(ops[i])
        for j in succs[i]:
            indeg[j] -= 1
            if indeg[j] == 0:
                ready.add(j)

    def packed(indices):
        # order commuting diagonal gates so that the kernel's sequential folding (consecutive gates while they touch at
        # most _RUN_BITS tile bits) yields few, long runs: grow one run at a time by the gate that adds the fewest bits
        rest, order = list(indices), []
        while rest:
            bits = set(qsets[rest[0]])
            order.append(rest.pop(0))
            while rest:
                best, best_new = None, 99
                for j in rest:
                    new = len(qsets[j] - bits)
                    if len(bits) + new <= _RUN_BITS and new < best_new:
                        best, best_new = j, new
                        if new == 0:
                            break
                if best is None:
                    break
                bits |= qsets[best]
                order.append(best)
                rest.remove(best)
        return order

    while ready:
        dense_ready = [i for i in sorted(ready) if not diag[i]]
        if not dense_ready:
            # diagonal gates nothing in the pass waits for (a QFT fan's phases onto qubits whose Hadamard comes in a later
            # pass) stay back: a later block may absorb them for free, and what is left at the end folds into long runs
            needed = [i for i in sorted(ready) if succs[i]]
            for i in packed(needed if needed else sorted(ready)):
                emit(i)
            continue
        seed = dense_ready[0]
        emit(seed)
        if ops[seed].kind != _lib.GATE_DENSE2:
            continue
        block = set(ops[seed].qubits)
        while True:
            pick = None
            for j in sorted(ready):
                if qsets[j] <= block or (len(block) == 2 and len(qsets[j]) == 2 and len(block | qsets[j]) == 3):
                    pick = j
                    break
            if pick is None:
                break
            block |= qsets[pick]
            emit(pick)
    assert len(out) == n
    return out


def pair_isolated_1q(ops):
    """Dense 1-qubit gates of a pass whose qubit no other gate of the pass touches (a layer of state-preparation
    rotations, the first Hadamard layer) are multiplied up two by two into 2-qubit gates M_b (x) M_a: the kernel widens
    a lone 1-qubit gate to a 4x4 block anyway (tile_pass.cu), so a pair costs one FP64 tensor op instead of two.  The
    gates commute with everything else in the pass; each pair takes the position of its first member."""
    touched = {}
    for op in ops:
        for q in op.qubits:
            touched[q] = touched.get(q, 0) + 1
    lone = [i for i, op in enumerate(ops) if op.kind == _lib.GATE_DENSE1 and touched[op.qubits[0]] == 1]
    if len(lone) < 2:
        return list(ops)
    out = list(ops)
    for a, b in zip(lone[0::2], lone[1::2]):
        qa, qb = ops[a].qubits[0], ops[b].qubits[0]
        mat = np.kron(np.asarray(ops[b].matrix, dtype=complex), np.asarray(ops[a].matrix, dtype=complex))
        out[a] = GateOp(_lib.GATE_DENSE2, [qa, qb], mat)
        out[b] = None
    return [op for op in out if op is not None]


_SWAP = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=complex)
MIN_SWAPS_TO_ELIMINATE = 4  # fewer swaps ride along in the passes of their neighbours for less than a pass of their own


def is_swap(op):
    """A dense 2-qubit op that is exactly the SWAP permutation matrix (three merged cx gates, qasm_simulator.py:145-161:
    the products of 0/1 matrices are exact)."""
    return op.kind == _lib.GATE_DENSE2 and op.matrix is not None and np.array_equal(np.asarray(op.matrix), _SWAP)


def eliminate_swaps(ops, n_qubits, low_qubits=None, min_swaps=MIN_SWAPS_TO_ELIMINATE):
    """Take the SWAP gates out of a (merged) op stream by relabelling: SWAP(a, b) G(a, ...) = G(b, ...) SWAP(a, b), so a
    swap moves to the end of the stream when the qubits of every later gate are renamed, and all swaps together leave ONE
    qubit permutation to settle there -- by qsv_swap_qubit_pairs, one in-place pass over HBM per round of disjoint
    transpositions, instead of a dense 4x4 tensor op per swap and a pass per three swaps (the final swaps of a QFT are
    16 swaps = 5 copy-only passes at 33 qubits; routing swaps of a transpiled circuit).

    Returns (ops without the eliminated swaps, rounds): ``rounds`` is a list of at most two lists of disjoint qubit pairs
    (any permutation is a product of two involutions); applying the ops, then round 0, then round 1 equals the input
    stream.  A swap on one of the ``low_qubits`` pinned bits (at the time it is met) stays a gate: the pass kernel moves
    those in 256-byte pieces, the permutation kernel would scatter 16-byte accesses.  Streams with fewer than
    ``min_swaps`` eliminable swaps are returned unchanged."""
    n_low = LOW_QUBITS if low_qubits is None else low_qubits
    where = list(range(n_qubits))  # where[q] = qubit that currently holds the content the stream calls q
    out, n_elim = [], 0
    for op in ops:
        if op.kind == "densek":
            return list(ops), []
        qs = [where[q] for q in op.qubits]
        if is_swap(op) and min(qs) >= n_low:
            where[op.qubits[0]], where[op.qubits[1]] = where[op.qubits[1]], where[op.qubits[0]]
            n_elim += 1
            continue
        out.append(op if list(op.qubits) == qs else GateOp(op.kind, qs, op.matrix))
    if n_elim < min_swaps:
        return list(ops), []
    # settle: the content of q sits at where[q] and has to move to q, i.e. position p -> m(p) with m = where^-1
    m = [0] * n_qubits
    for q, p in enumerate(where):
        m[p] = q
    rounds = [[], []]
    seen = [False] * n_qubits
    for start in range(n_qubits):
        if seen[start] or m[start] == start:
            continue
        cyc = []
        p = start
        while not seen[p]:
            seen[p] = True
            cyc.append(p)
            p = m[p]
        length = len(cyc)
        # m = t2 o t1 on the cycle c_0 -> c_1 -> ...: t1(c_i) = c_(-i), t2(c_i) = c_(1-i), two reflections
        for i in range(length):
            j = (-i) % length
            if i < j:
                rounds[0].append((cyc[i], cyc[j]))
            j = (1 - i) % length
            if i < j:
                rounds[1].append((cyc[i], cyc[j]))
    rounds = [r for r in rounds if r]
    if n_elim < min_swaps * len(rounds):  # chains of swaps (cycles) need two launches: they have to replace more gates
        return list(ops), []
    return out, rounds


def _relabel(op, layout):
    return GateOp(op.kind, [layout[q] for q in op.qubits], op.matrix)


def plan_passes(ops, n_qubits, tile_qubits=DEFAULT_TILE_QUBITS, max_gates=_lib.MAX_PASS_GATES,
                max_payload=_lib.MAX_PASS_MATRIX_DOUBLES, layout=None, remap=False, tails=False, schedule=True,
                avoid=(), must=None, leftover=None, pair_1q=True):
    """Group ``ops`` (GateOps on LOGICAL qubits, <= 2 qubits each) into passes.

    Frontier-greedy over the gate dependency DAG: a gate is *ready* when it is the first pending gate
    on each of its qubits; among the ready gates that still fit the tile, the one that adds the fewest
    new tile qubits is taken next (ties: program order), until nothing fits.  Only gates on disjoint
    qubits are ever commuted.  (On QuantumVolume(33) this yields 7.3 gates per 11-qubit pass against
    5.7 for a plain program-order scan.)

    With ``tails`` every pass also takes a diagonal TAIL (``Pass.tail``, qsv_apply_pass_tail): once the
    pass's gates are chosen, every diagonal gate that is next in line on all of its qubits and has a qubit
    outside the tile (or is a 1-qubit diagonal) rides along -- the kernel applies the whole tail with one or
    two table look-ups per amplitude, wherever the gates' qubits are.  A controlled-phase ladder (QFT) then
    needs a pass per ~7 Hadamards instead of a pass per handful of controlled phases.

    ``layout[q]`` is the physical bit of logical qubit q (identity if None); it is updated in place when
    given.  With ``remap`` the write-back of each pass moves the tile qubits that the following gates
    need first into the always-resident low bits (qsv_apply_pass_remap).

    ``must`` (a set of indices into ``ops``) turns the call into a PARTIAL plan (dist.ShardedState before a qubit
    exchange): only the ``must`` gates have to run now.  They are preferred whenever one fits; other ready gates only
    fill the tiles the ``must`` gates open, and planning stops with the pass that takes the last ``must`` gate.  The
    indices of the gates left unplanned (closed under successors, program order) are appended to ``leftover``."""
    from collections import deque

    if layout is None:
        layout = list(range(n_qubits))
    tile_size = min(n_qubits, tile_qubits, _lib.MAX_TILE_QUBITS)
    n_low = min(LOW_QUBITS, n_qubits)
    for op in ops:
        if len(op.qubits) > 2 or op.kind == "densek":
            raise ValueError("plan_passes only fuses 1- and 2-qubit gates")
    if tile_size < n_low + 2:
        n_low = 0  # tiny tiles (tests): no pinned low bits
    remap = remap and n_qubits > tile_size and n_low > 0

    avoid_set = set(avoid)
    queues = {}
    for idx, op in enumerate(ops):
        for q in op.qubits:
            queues.setdefault(q, deque()).append(idx)

    def is_ready(idx):
        return all(queues[q][0] == idx for q in ops[idx].qubits)

    frontier = set(idx for idx in range(len(ops)) if is_ready(idx))
    n_left = len(ops)
    passes = []
    must_left = None if must is None else len(set(must))
    taken = set()
    while n_left and (must is None or must_left > 0):
        inv = [0] * n_qubits
        for q, ph in enumerate(layout):
            inv[ph] = q
        low_log = [inv[ph] for ph in range(n_low)]  # logical qubits sitting in the pinned low bits
        used = set(low_log)
        chosen = []
        payload = 0
        while len(chosen) < max_gates:
            best, best_cost = None, 99
            for idx in frontier:
                op = ops[idx]
                cost = sum(1 for q in op.qubits if q not in used)
                if len(used) + cost > tile_size or payload + op.payload_doubles > max_payload:
                    continue
                if must is not None and idx not in must:
                    cost += 50  # a gate that may wait only fills what the gates that must run leave free
                if cost < best_cost or (cost == best_cost and idx < best):
                    best, best_cost = idx, cost
            if best is None:
                break
            op = ops[best]
            frontier.discard(best)
            taken.add(best)
            chosen.append(op)
            used.update(op.qubits)
            payload += op.payload_doubles
            n_left -= 1
            for q in op.qubits:
                queues[q].popleft()
            for q in op.qubits:
                if queues[q] and is_ready(queues[q][0]):
                    frontier.add(queues[q][0])
        if not chosen:  # cannot happen: a ready gate always fits an empty tile
            raise RuntimeError("planner made no progress")
        # next pending gate of every qubit (position in program order)
        next_use = {q: dq[0] for q, dq in queues.items() if dq}
        far = len(ops) + 1
        # fill the tile: qubits needed soonest first (they can be made resident), then the lowest bits
        if len(used) < tile_size:
            # ``avoid``: qubits that must stay outside the tile when no gate needs them (the fixed qubits of a
            # sub-block pass, qsv_apply_pass_sub)
            cand = sorted((q for q in range(n_qubits) if q not in used and q not in avoid_set),
                          key=lambda q: (next_use.get(q, far), layout[q]))
            used |= set(cand[: tile_size - len(used)])
        tile_log = sorted(used, key=lambda q: layout[q])
        tile_phys = [layout[q] for q in tile_log]
        tail = []
        if tails:
            diag = (_lib.GATE_DIAG1, _lib.GATE_DIAG2)

            def take(idx):
                nonlocal n_left
                frontier.discard(idx)
                taken.add(idx)
                n_left -= 1
                for q in ops[idx].qubits:
                    queues[q].popleft()
                for q in ops[idx].qubits:
                    if queues[q] and is_ready(queues[q][0]):
                        frontier.add(queues[q][0])

            grew = True
            while grew:
                grew = False
                for idx in sorted(frontier):
                    op = ops[idx]
                    if op.kind not in diag:
                        continue
                    if len(op.qubits) == 2 and all(q in used for q in op.qubits):
                        # a 2-qubit diagonal inside the tile is an ordinary gate of the pass.  One that became ready
                        # only behind tail gates (the controlled phases onto the pinned low qubits at the end of a QFT
                        # fan) still joins THIS pass: it is appended to the pass's gates, which run before the tail --
                        # diagonal gates commute with the (all-diagonal) tail, and every non-diagonal gate before it on
                        # its qubits is already in the pass or done
                        if not tail or len(chosen) >= max_gates or payload + op.payload_doubles > max_payload:
                            continue
                        chosen.append(op)
                        payload += op.payload_doubles
                        take(idx)
                        grew = True
                        continue
                    if len(tail) >= _lib.MAX_TAIL_GATES:
                        continue
                    tail.append(op)
                    take(idx)
                    grew = True
        dest = None
        if remap and n_left:
            # the n_low tile qubits needed soonest become the residents of the low bits
            order = sorted(tile_log, key=lambda q: (next_use.get(q, far), layout[q] >= n_low, layout[q]))
            sticky = set(order[:n_low])
            incoming = [q for q in tile_log if q in sticky and layout[q] >= n_low]
            evicted = [q for q in low_log if q not in sticky]
            assert len(incoming) == len(evicted)
            new_pos = {q: layout[q] for q in tile_log}
            for qi, qe in zip(incoming, evicted):
                new_pos[qi], new_pos[qe] = layout[qe], layout[qi]
            if incoming:
                dest = [new_pos[q] for q in tile_log]
        if pair_1q:
            chosen = pair_isolated_1q(chosen)
        if schedule:
            chosen = schedule_pass_ops(chosen)
        passes.append(Pass(tile_phys, [_relabel(op, layout) for op in chosen], dest,
                           [_relabel(op, layout) for op in tail]))
        if dest is not None:
            for q, d in zip(tile_log, dest):
                layout[q] = d
        if must is not None:
            must_left = sum(1 for idx in must if idx not in taken)
    if leftover is not None:
        leftover.extend(idx for idx in range(len(ops)) if idx not in taken)
    return passes


def plan_restore(layout, n_qubits, tile_qubits=DEFAULT_TILE_QUBITS):
    """Passes (no gates, only write-back permutations) that bring ``layout`` back to the identity, i.e.
    every logical qubit q to physical bit q.  ``layout`` is updated in place."""
    tile_size = min(n_qubits, tile_qubits, _lib.MAX_TILE_QUBITS)
    n_low = min(LOW_QUBITS, n_qubits)
    if tile_size < n_low + 2:
        n_low = 0
    passes = []
    guard = 0
    while any(layout[q] != q for q in range(n_qubits)):
        guard += 1
        if guard > 4 * n_qubits:
            raise RuntimeError("plan_restore does not converge")
        inv = [0] * n_qubits
        for q, ph in enumerate(layout):
            inv[ph] = q  # content of position ph is logical qubit inv[ph]; its home is position inv[ph]
        members = list(range(n_low))
        in_tile = set(members)
        dest = {}
        starts = [ph for ph in range(n_qubits) if inv[ph] != ph]
        starts.sort(key=lambda ph: (ph not in in_tile, ph))
        for start in starts:
            if start in dest:
                continue
            if start not in in_tile and len(in_tile) + 2 > tile_size:
                continue
            chain = [start]
            extra = 0 if start in in_tile else 1
            nxt = inv[start]
            closed = False
            while True:
                if nxt == start:
                    closed = True
                    break
                if nxt in dest or nxt in chain:
                    break
                cost = 0 if nxt in in_tile else 1
                if len(in_tile) + extra + cost > tile_size:
                    break
                chain.append(nxt)
                extra += cost
                nxt = inv[nxt]
            if len(chain) < 2:
                continue
            for i, ph in enumerate(chain):
                if i + 1 < len(chain):
                    dest[ph] = chain[i + 1]
                else:
                    dest[ph] = start if not closed else inv[ph]
            for ph in chain:
                if ph not in in_tile:
                    in_tile.add(ph)
                    members.append(ph)
        if not dest:
            raise RuntimeError("plan_restore made no progress")
        for ph in members:
            dest.setdefault(ph, ph)
        # fill the tile up to its size with untouched positions (keeps the kernel's tile shape uniform)
        ph = 0
        while len(members) < tile_size and ph < n_qubits:
            if ph not in in_tile:
                in_tile.add(ph)
                members.append(ph)
                dest[ph] = ph
            ph += 1
        members.sort()
        passes.append(Pass(members, [], [dest[ph] for ph in members]))
        for ph in members:
            layout[inv[ph]] = dest[ph]
    return passes
