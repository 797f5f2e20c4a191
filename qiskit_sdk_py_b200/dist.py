"""Multi-GPU statevector: one process per GPU, the state sharded by its high-order qubits.

The reference has no multi-device path (qasm_simulator.py keeps one numpy array in host memory, capped
at 24 qubits); this is the B200 extension BASELINE config 5 asks for.  Layout:

  * W = 2^g ranks; rank r holds the 2^(n-g) amplitudes whose top g *physical* index bits equal r.
  * ``pos[q]`` maps logical qubit q to its physical bit.  Physical bits < n_local are local, the others
    are "global" (rank bits).
  * Gates whose non-diagonal action touches only local bits run with no communication, on every rank,
    through the same fused pass kernel as the single-GPU path.  Diagonal gates and CX controls on a
    global qubit also need no communication (the rank's bit selects a factor / enables the X).
  * A dense gate (or a CX target) on a global qubit triggers a **qubit swap**: global bit j <-> local
    bit L.  Each rank exchanges, with partner r ^ 2^j, the half of its shard whose bit L differs from
    its own bit j -- pack (qsv_pack_half) -> NCCL send/recv over NVLink -> unpack (qsv_unpack_half),
    chunked through two staging buffers.  The victim L is the local qubit whose next use lies furthest
    in the future.
  * Reductions (P(qubit), norm, sampler totals) are all-reduces / all-gathers of a few doubles.

The transport is ``torch.distributed`` (NCCL on GPUs; gloo in the CPU tests, where the shard is the numpy
test double) -- plumbing only; every operation on amplitudes is a libqsv kernel.
"""
import math
import os

import numpy as np

from . import _lib
from .lower import GateOp
from .planner import DEFAULT_TILE_QUBITS as DEFAULT_TILE

DEFAULT_CHUNK_AMPS = 1 << 26  # 1 GiB staging buffers (torch.distributed transport of the CPU tests)
_IPC_OPEN = {}                 # CUDA IPC handle -> [mapped base pointer, references]
_LOW_BITS = 4                  # the planner pins the lowest local bits into every tile (256-byte runs)


class ShardedState:
    """A 2^n statevector sharded over ``world`` ranks.  ``shard`` is an engine.DeviceState (or a test
    double with the same interface) holding this rank's 2^(n - log2(world)) amplitudes."""

    def __init__(self, n_qubits, shard_factory, rank=None, world=None, chunk_amps=DEFAULT_CHUNK_AMPS,
                 group=None, p2p=True, overlap=True):
        import torch
        import torch.distributed as dist

        self._torch = torch
        self._dist = dist
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.g = int(round(math.log2(self.world)))
        if (1 << self.g) != self.world:
            raise ValueError("world size must be a power of two")
        self.n = int(n_qubits)
        self.n_local = self.n - self.g
        if self.n_local < 2:
            raise ValueError("need at least two local qubits per rank")
        self.shard = shard_factory(self.n_local)
        self.pos = list(range(self.n))          # logical qubit -> physical bit
        self.chunk_amps = int(chunk_amps)
        self._send = None
        self._recv = None
        self.swaps = 0
        self.fused_exchanges = 0
        self.swap_bytes = 0
        self._xch_events = []   # (start, end) CUDA events around every exchange (device shards only)
        self._xch_wait_events = []  # (start, after the opening barrier) of the exchanges that are not overlapped
        self.overlap = bool(overlap)  # run the passes that commute with an exchange while it is in flight
        self.overlapped_exchanges = 0
        self.relabelled_swaps = 0
        self._comm_stream = None
        # NVLink peer memory: shards that can export a CUDA IPC handle (engine.DeviceState) are mapped into every
        # rank of the group once, and the exchanges run as ONE in-place swap kernel per rank (qsv_swap_bits_p2p);
        # the numpy test double of the CPU tests has no such handle and goes through torch.distributed send/recv
        self._peers = None      # group rank -> (peer pointer, offset)
        if p2p and hasattr(self.shard, "ipc_export"):
            self._map_peers()

    # -- helpers -----------------------------------------------------------------------
    def _logical_at(self, phys):
        return self.pos.index(phys)

    def _rank_bit(self, phys):
        return (self.rank >> (phys - self.n_local)) & 1

    def _global_rank(self, group_rank):
        """Rank inside ``self.group`` -> rank in the default (world) group."""
        if self.group is None:
            return group_rank
        return self._dist.get_global_rank(self.group, group_rank)

    def _staging(self, amps):
        """Two send and two receive buffers of ``amps`` amplitudes (the exchange is double-buffered)."""
        torch = self._torch
        dev = self.shard.comm_device()
        if self._send is None or self._send[0].numel() < 2 * amps:
            self._send = [torch.empty(2 * amps, dtype=torch.float64, device=dev) for _ in range(2)]
            self._recv = [torch.empty(2 * amps, dtype=torch.float64, device=dev) for _ in range(2)]
        return self._send, self._recv

    # -- initialisation ----------------------------------------------------------------
    def init_basis0(self, phase=1.0 + 0.0j):
        self.pos = list(range(self.n))
        if self.rank == 0:
            self.shard.init_basis0(phase)
        else:
            self.shard.init_basis0(0.0 + 0.0j)

    def init_identity(self, n_half, phase=1.0 + 0.0j):
        """The identity matrix as a vector of 2 n_half qubits (row index = the high half; unitary simulator,
        unitary_simulator.py:190-199), built with exact arithmetic from |0..0>: the map |0> -> |0> + |1> on every
        row qubit (entries 0/1, not a Hadamard), then CX row_j -> column_j gives sum_k |k>|k>."""
        if 2 * n_half != self.n:
            raise _lib.QsvLibraryError("init_identity: the state has %d qubits, not 2 * %d" % (self.n, n_half))
        self.init_basis0(phase)
        fan = np.array([[1, 0], [1, 0]], dtype=complex)
        self.apply_ops([GateOp(_lib.GATE_DENSE1, [n_half + j], fan) for j in range(n_half)] +
                       [GateOp(_lib.GATE_CX, [n_half + j, j], None) for j in range(n_half)])

    # -- NVLink peer mapping -------------------------------------------------------------------------
    def _map_peers(self):
        handle, offset = self.shard.ipc_export()
        everyone = [None] * self.world
        self._dist.all_gather_object(everyone, (handle, offset), group=self.group)
        self._peers = {}
        for r, (h, off) in enumerate(everyone):
            if r == self.rank:
                continue
            # an allocation can be opened only once per process: successive states of a peer usually live in the same
            # cached block, so the mapping is shared and reference-counted
            entry = _IPC_OPEN.get(h)
            if entry is None:
                entry = _IPC_OPEN[h] = [self.shard.ipc_import(h, 0), 0]
            entry[1] += 1
            self._peers[r] = (entry[0] + off, h)

    def _unmap_peers(self):
        peers, self._peers = self._peers, None
        for _ptr, h in (peers or {}).values():
            entry = _IPC_OPEN.get(h)
            if entry is None:
                continue
            entry[1] -= 1
            if entry[1] <= 0:
                del _IPC_OPEN[h]
                self.shard.ipc_release(entry[0], 0)

    def close(self):
        """Unmap the peers' shards (collective: every rank calls it before any rank frees its shard)."""
        if self._peers:
            self._torch.cuda.synchronize()
            self._unmap_peers()
            self._dist.barrier(group=self.group)

    def __del__(self):
        try:
            self._unmap_peers()
        except Exception:  # interpreter shutdown
            pass

    def _stream_barrier(self):
        """Stream-ordered barrier among the ranks: everything enqueued before it has completed on EVERY rank before
        anything enqueued after it starts (a one-element all-reduce on the compute stream)."""
        t = self._torch.zeros(1, dtype=self._torch.float32, device=self.shard.comm_device())
        self._dist.all_reduce(t, group=self.group)

    def _swap_p2p(self, pairs):
        """The exchange as one in-place NVLink kernel per rank (see qsv_swap_bits_p2p)."""
        k = len(pairs)
        if k > 1 and self.n_local - k < 1:
            # tiny shards (2 local qubits on 8 ranks): a sub-block of one amplitude has no halves for the two partners
            # to share -- the same permutation as k pairwise swaps (the pairs are disjoint)
            for pair in pairs:
                self._swap_p2p([pair])
            return
        t_begin = self._xch_begin()
        gbits = [pg - self.n_local for pg, _ in pairs]
        lbits = [pl for _, pl in pairs]
        beta = 0
        for i, j in enumerate(gbits):
            beta |= ((self.rank >> j) & 1) << i
        ptrs = [None] * (1 << k)
        for lam in range(1 << k):
            if lam == beta:
                continue
            r = self.rank
            for i, j in enumerate(gbits):
                r = (r & ~(1 << j)) | (((lam >> i) & 1) << j)
            ptrs[lam] = self._peers[r][0]
        self._stream_barrier()
        if t_begin is not None:  # time spent waiting for the slowest rank (bench.py: exchange_wait_ms)
            mid = self._torch.cuda.Event(enable_timing=True)
            mid.record()
            self._xch_wait_events.append((t_begin, mid))
        if k > 1 and self.p2p_stepwise:
            # one launch per partner (XOR order) with a barrier in between: every step is a perfect matching of the ranks;
            # inside ONE kernel the ranks drift apart and two of them end up loading the same partner's links
            for j in range(1, 1 << k):
                if j > 1:
                    self._stream_barrier()
                step = [None] * (1 << k)
                step[beta ^ j] = ptrs[beta ^ j]
                self.shard.swap_bits_p2p(lbits, beta, step)
        else:
            self.shard.swap_bits_p2p(lbits, beta, ptrs)
        self._stream_barrier()
        self._xch_end(t_begin)
        for pg, pl in pairs:
            qa, qb = self._logical_at(pg), self._logical_at(pl)
            self.pos[qa], self.pos[qb] = pl, pg
        self.swaps += k
        if k > 1:
            self.fused_exchanges += 1
        self.swap_bytes += 16 * (1 << (self.n_local - k)) * ((1 << k) - 1)

    # -- exchange timing (bench.py: exchange vs pass time per step) -------------------------------
    def _xch_begin(self):
        if getattr(self.shard, "device", None) is None:
            return None
        ev = self._torch.cuda.Event(enable_timing=True)
        ev.record()
        return ev

    def _xch_end(self, begin):
        if begin is not None:
            ev = self._torch.cuda.Event(enable_timing=True)
            ev.record()
            self._xch_events.append((begin, ev))

    @property
    def exchange_ms(self):
        """Marker for exchange_time_ms: exchanges recorded so far."""
        return len(self._xch_events)

    def exchange_wait_ms(self):
        """Device time (ms) this rank spent in the opening barrier of its (not overlapped) exchanges so far, i.e. waiting
        for the slowest rank (synchronises)."""
        self._torch.cuda.synchronize()
        return float(sum(a.elapsed_time(b) for a, b in self._xch_wait_events))

    def exchange_time_ms(self, since=0):
        """Device time (ms) of the exchanges recorded after marker ``since`` (synchronises)."""
        if not self._xch_events[since:]:
            return 0.0
        self._torch.cuda.synchronize()
        return float(sum(a.elapsed_time(b) for a, b in self._xch_events[since:]))

    # -- the exchange step -------------------------------------------------------------
    def swap_global_local(self, phys_global, phys_local):
        """Swap physical bits (global j, local L): exchange the half with bit L != my bit j."""
        if self._peers is not None:
            return self._swap_p2p([(phys_global, phys_local)])
        dist = self._dist
        t_begin = self._xch_begin()
        j = phys_global - self.n_local
        beta = (self.rank >> j) & 1
        peer = self.rank ^ (1 << j)
        half = 1 << (self.n_local - 1)
        chunk = min(self.chunk_amps, half)
        send, recv = self._staging(chunk)
        spans = [(begin, min(chunk, half - begin)) for begin in range(0, half, chunk)]

        gpeer = self._global_rank(peer)  # P2POp takes GLOBAL ranks also when a group is given

        def exchange(i):
            cnt = spans[i][1]
            ops = [dist.P2POp(dist.isend, send[i % 2][: 2 * cnt], gpeer, self.group),
                   dist.P2POp(dist.irecv, recv[i % 2][: 2 * cnt], gpeer, self.group)]
            if self.rank > peer:
                ops.reverse()
            return dist.batch_isend_irecv(ops)

        # software pipeline: the NCCL exchange of chunk i+1 runs while chunk i is unpacked and chunk i+2
        # is packed (the collective only waits for work enqueued before it was issued)
        self.shard.pack_half(phys_local, 1 - beta, send[0], spans[0][0], spans[0][1])
        reqs = exchange(0)
        if len(spans) > 1:
            self.shard.pack_half(phys_local, 1 - beta, send[1], spans[1][0], spans[1][1])
        for i, (begin, cnt) in enumerate(spans):
            for req in reqs:
                req.wait()
            if i + 1 < len(spans):
                reqs = exchange(i + 1)
            self.shard.unpack_half(phys_local, 1 - beta, recv[i % 2], begin, cnt)
            if i + 2 < len(spans):
                self.shard.pack_half(phys_local, 1 - beta, send[i % 2], spans[i + 2][0], spans[i + 2][1])
        self._xch_end(t_begin)
        qa, qb = self._logical_at(phys_global), self._logical_at(phys_local)
        self.pos[qa], self.pos[qb] = phys_local, phys_global
        self.swaps += 1
        self.swap_bytes += 16 * half

    def swap_many(self, pairs):
        """Swap k global bits with k local bits in ONE all-to-all among the 2^k ranks that differ only in
        those global bits: ``pairs`` = [(phys_global, phys_local), ...].  Each rank keeps 1/2^k of its shard
        and exchanges the other sub-blocks, (1 - 2^-k) of the shard in total -- against k/2 of the shard for
        k successive pairwise swaps."""
        if self._peers is not None:
            return self._swap_p2p(pairs)
        if len(pairs) == 1:
            return self.swap_global_local(*pairs[0])
        t_begin = self._xch_begin()
        k = len(pairs)
        gbits = [pg - self.n_local for pg, _ in pairs]
        lbits = [pl for _, pl in pairs]
        beta = 0
        for i, j in enumerate(gbits):
            beta |= ((self.rank >> j) & 1) << i
        peers = {}
        for lam in range(1 << k):
            if lam == beta:
                continue
            r = self.rank
            for i, j in enumerate(gbits):
                r = (r & ~(1 << j)) | (((lam >> i) & 1) << j)
            peers[lam] = r
        self._swap_staged(lbits, beta, peers)
        self._xch_end(t_begin)
        for pg, pl in pairs:
            qa, qb = self._logical_at(pg), self._logical_at(pl)
            self.pos[qa], self.pos[qb] = pl, pg
        self.swaps += k
        self.fused_exchanges += 1
        self.swap_bytes += 16 * (1 << (self.n_local - k)) * ((1 << k) - 1)

    def _swap_staged(self, lbits, beta, peers):
        """torch.distributed transport (CPU tests): my sub-blocks [lbits = lam] trade places with rank peers[lam]'s
        sub-block [lbits = beta], for every lam in ``peers``; pack -> send/recv -> unpack through staging buffers."""
        dist = self._dist
        k = len(lbits)
        sub = 1 << (self.n_local - k)
        chunk = max(1, min(self.chunk_amps // max(1, len(peers)), sub))
        spans = [(begin, min(chunk, sub - begin)) for begin in range(0, sub, chunk)]
        send, recv = self._staging(chunk * len(peers))
        lams = sorted(peers)

        def view(bufs, i, slot, cnt):
            return bufs[i % 2][2 * slot * chunk: 2 * slot * chunk + 2 * cnt]

        def pack(i):
            begin, cnt = spans[i]
            for slot, lam in enumerate(lams):
                self.shard.pack_bits(lbits, lam, view(send, i, slot, cnt), begin, cnt)

        def exchange(i):
            cnt = spans[i][1]
            ops = []
            for slot, lam in enumerate(lams):
                gpeer = self._global_rank(peers[lam])  # P2POp takes GLOBAL ranks also when a group is given
                ops.append(dist.P2POp(dist.isend, view(send, i, slot, cnt), gpeer, self.group))
                ops.append(dist.P2POp(dist.irecv, view(recv, i, slot, cnt), gpeer, self.group))
            return dist.batch_isend_irecv(ops)

        def unpack(i):
            begin, cnt = spans[i]
            for slot, lam in enumerate(lams):
                # what arrives from the rank whose global bits are lam lands where my bits lam were
                self.shard.unpack_bits(lbits, lam, view(recv, i, slot, cnt), begin, cnt)

        pack(0)
        reqs = exchange(0)
        if len(spans) > 1:
            pack(1)
        for i in range(len(spans)):
            for req in reqs:
                req.wait()
            if i + 1 < len(spans):
                reqs = exchange(i + 1)
            unpack(i)
            if i + 2 < len(spans):
                pack(i + 2)

    # -- gates -------------------------------------------------------------------------
    def _localise(self, op):
        """Return the op rewritten on physical local bits, ``None`` if it is a no-op on this rank, or
        ``False`` if it needs a qubit swap first."""
        phys = [self.pos[q] for q in op.qubits]
        is_global = [p >= self.n_local for p in phys]
        if not any(is_global):
            return GateOp(op.kind, phys, op.matrix)
        if op.kind == _lib.GATE_DIAG1:
            f = op.matrix[self._rank_bit(phys[0])]
            return ("scale", complex(f))
        if op.kind == _lib.GATE_DIAG2:
            d = np.asarray(op.matrix).reshape(2, 2)  # d[bit1][bit0]
            if all(is_global):
                return ("scale", complex(d[self._rank_bit(phys[1])][self._rank_bit(phys[0])]))
            if is_global[0]:
                return GateOp(_lib.GATE_DIAG1, [phys[1]], d[:, self._rank_bit(phys[0])].copy())
            return GateOp(_lib.GATE_DIAG1, [phys[0]], d[self._rank_bit(phys[1]), :].copy())
        if op.kind == _lib.GATE_CX and not is_global[1]:
            if self._rank_bit(phys[0]):
                return GateOp(_lib.GATE_DENSE1, [phys[1]], np.array([[0, 1], [1, 0]], dtype=complex))
            return None
        return False

    def apply_ops(self, ops):
        """Apply GateOps (logical qubits, program order).  Gates that need no exchange are fused into
        local passes -- commuting over deferred gates on disjoint qubits only -- then the global qubits
        the deferred gates need are swapped in, and so on."""
        ops = list(ops)
        for k, op in enumerate(ops):
            if op.kind == "densek":  # a k >= 3 qubit unitary is a barrier: run what precedes it, then it
                if len(op.qubits) > self.n_local:
                    raise _lib.QsvLibraryError("a %d-qubit unitary does not fit in a shard of %d qubits"
                                               % (len(op.qubits), self.n_local))
                self.apply_ops(ops[:k])
                self._apply_dense_k(op)
                self.apply_ops(ops[k + 1:])
                return
        remaining = self._mark_swaps(ops)
        while remaining:
            run, deferred, blocked = [], [], set()
            plain = True   # every runnable op acts on local qubits only: the same GateOps on every rank
            for op in remaining:
                qs = set(op.qubits)
                if op.kind == "swap":
                    # a SWAP gate is a relabelling of the two logical qubits: no data moves now, the layout restore (or
                    # the next exchange that needs one of them) pays for it -- on any rank, global qubits included
                    if qs & blocked:
                        blocked |= qs
                        deferred.append(op)
                    else:
                        a, b = op.qubits
                        self.pos[a], self.pos[b] = self.pos[b], self.pos[a]
                        self.relabelled_swaps += 1
                    continue
                loc = False if (qs & blocked) else self._localise(op)
                if loc is False:
                    blocked |= qs
                    deferred.append(op)
                elif loc is not None:
                    run.append(loc)
                if loc is not False and any(self.pos[q] >= self.n_local for q in op.qubits):
                    plain = False
            pairs = self._plan_swap(deferred) if deferred else []
            late, carry = [], []
            # with k swapped qubits a fraction 1 - 2^-k of the late work runs under the transfers: worth the extra
            # (less fused) passes from k = 2 on; a pairwise swap (k = 1) hides too little (measured on 2 GPUs)
            want_late = len(pairs) >= self.MIN_OVERLAP_SWAPS and self.overlap
            if pairs and plain and self.carry_over:
                late, carry = self._run_must_then_split(run, [pl for _, pl in pairs], want_late)
            else:
                if want_late:
                    run, late = self._split_for_overlap(run, [pl for _, pl in pairs])
                self._apply_local(run)
            if late:
                self._exchange_overlapped(pairs, late)
            elif pairs:
                self.swap_many(pairs)
            remaining = carry + deferred

    @staticmethod
    def _mark_swaps(ops):
        """cx a,b; cx b,a; cx a,b (a SWAP in the simulator basis, back to back) -> one ("swap", [a, b]) marker op."""
        out, i = [], 0
        while i < len(ops):
            o = ops[i]
            if (o.kind == _lib.GATE_CX and i + 2 < len(ops) and ops[i + 1].kind == _lib.GATE_CX
                    and ops[i + 2].kind == _lib.GATE_CX and tuple(ops[i + 1].qubits) == tuple(o.qubits)[::-1]
                    and tuple(ops[i + 2].qubits) == tuple(o.qubits)):
                out.append(GateOp("swap", list(o.qubits), None))
                i += 3
            else:
                out.append(o)
                i += 1
        return out

    p2p_stepwise = True   # fused peer-memory exchanges partner by partner (4 GPUs: 1886 -> 1815 ms of exchanges per step)
    carry_over = True  # gates that need not run before an exchange wait for the passes after it (fuller passes)

    def _run_must_then_split(self, run, victim_bits, want_late):
        """Before an exchange only the gates on a victim bit (the local bits about to leave the shard) and their
        predecessors MUST run.  Plan exactly those (planner.plan_passes(must=...): other ready gates just fill the tiles
        these open), apply the passes, and split what is left -- gates on other qubits, closed under successors -- into
        ``late`` (a prefix of at most OVERLAP_TARGET_GATES gates, applied sub-block by sub-block under the transfers)
        and ``carry`` (returned as ops on LOGICAL qubits: they join the gates that become runnable after the exchange,
        where they fuse into full passes instead of the sparse last passes of this batch).  ``run`` holds plain local
        GateOps, the same on every rank, so every rank takes the same decisions."""
        from .planner import merge_ops, plan_passes

        merged = merge_ops(run)
        tainted = set(victim_bits)
        must = set()
        for idx in range(len(merged) - 1, -1, -1):
            qs = set(merged[idx].qubits)
            if qs & tainted:
                tainted |= qs
                must.add(idx)
        left = []
        if must:
            for p in plan_passes(merged, self.n_local, self.shard.tile_qubits, tails=True, must=must, leftover=left):
                self.shard.apply_pass(p.tile_qubits, p.ops, tail=p.tail)
        else:
            left = list(range(len(merged)))
        rest = [merged[idx] for idx in left]
        n_late = 0
        if want_late and self.n_local - len(victim_bits) >= self.shard.tile_qubits and \
                not any(v < _LOW_BITS for v in victim_bits):
            n_late = min(len(rest), self.OVERLAP_TARGET_GATES)
            if n_late < self.MIN_OVERLAP_GATES:
                n_late = 0
        late = rest[:n_late]
        carry = [GateOp(op.kind, [self._logical_at(b) for b in op.qubits], op.matrix) for op in rest[n_late:]]
        return late, carry

    # -- exchange overlapped with the passes that commute with it ---------------------------------------
    MIN_OVERLAP_SWAPS = 2    # overlap only fused exchanges of at least this many qubits
    MIN_OVERLAP_GATES = 6    # fewer late gates than this: not worth splitting the passes
    OVERLAP_TARGET_GATES = 30  # ... and no more than about four passes' worth (one exchange takes 2-4 pass times)
    EXCHANGE_SMS = 32        # SMs left to the exchange kernel while sub-block passes run (4 GPUs: 16 -> 6.90 s, 32 -> 6.72 s)

    def _split_for_overlap(self, run, victim_bits):
        """Split the locally runnable ops (physical qubits, program order) into (early, late): ``late`` = the ops
        with no successor-or-self on a victim bit -- they act on other qubits than the exchange moves, and nothing
        that touches a victim bit has to wait for them, so they commute with the exchange and can run sub-block by
        sub-block while it is in flight.  Global scalars stay early."""
        if self.n_local - len(victim_bits) < self.shard.tile_qubits or any(v < _LOW_BITS for v in victim_bits):
            return run, []
        tainted = set(victim_bits)
        late_flags = [False] * len(run)
        for idx in range(len(run) - 1, -1, -1):
            item = run[idx]
            if isinstance(item, tuple):
                continue
            if set(item.qubits) & tainted:
                tainted |= set(item.qubits)
            else:
                late_flags[idx] = True
        # no more late work than the exchange can hide (the last OVERLAP_TARGET_GATES of them: a suffix of the late ops
        # is closed under successors); the rest stays with the early ops, where it fuses into fuller passes
        n_late = 0
        for idx in range(len(run) - 1, -1, -1):
            if late_flags[idx]:
                n_late += 1
                if n_late > self.OVERLAP_TARGET_GATES:
                    late_flags[idx] = False
        late = [op for op, f in zip(run, late_flags) if f]
        if len(late) < self.MIN_OVERLAP_GATES:
            return run, []
        return [op for op, f in zip(run, late_flags) if not f], late

    def _exchange_overlapped(self, pairs, late):
        """The exchange of ``pairs`` with the ``late`` ops applied meanwhile: the swapped local bits split the shard
        into 2^k sub-blocks; the late passes run on one sub-block after the other (qsv_apply_pass_sub, the swapped
        bits fixed), and as soon as a sub-block is done on both partners it is exchanged (XOR schedule: in step j
        rank beta trades with beta ^ j) -- on the GPU by the peer-memory swap kernel on a second stream and on the
        SMs the passes leave free; the sub-block that stays is processed last, under the last transfers."""
        k = len(pairs)
        gbits = [pg - self.n_local for pg, _ in pairs]
        lbits = [pl for _, pl in pairs]
        beta = 0
        for i, j in enumerate(gbits):
            beta |= ((self.rank >> j) & 1) << i
        rank_of = {}
        for lam in range(1 << k):
            r = self.rank
            for i, j in enumerate(gbits):
                r = (r & ~(1 << j)) | (((lam >> i) & 1) << j)
            rank_of[lam] = r
        passes = self.shard.plan_ops(late, avoid=lbits)
        for p in passes:
            if set(p.tile_qubits) & set(lbits):
                raise _lib.QsvLibraryError("overlapped exchange: a late pass has a swapped bit in its tile")
        order = [beta ^ j for j in range(1, 1 << k)] + [beta]
        t_begin = self._xch_begin()
        device = self._peers is not None
        if device:
            torch = self._torch
            if self._comm_stream is None:
                self._comm_stream = torch.cuda.Stream()
            cur, comm = torch.cuda.current_stream(), self._comm_stream
            n_sms = self.shard.sm_count
            cap = max(1, n_sms - self.EXCHANGE_SMS)
        for lam in order:
            for p in passes:
                self.shard.apply_pass_sub(p.tile_qubits, p.ops, p.tail, lbits, lam, cap if device else 0)
            if lam == beta:
                continue
            if device:
                ev = torch.cuda.Event()
                ev.record(cur)
                with torch.cuda.stream(comm):
                    comm.wait_event(ev)
                    self._stream_barrier()  # this step's sub-block is finished on every rank
                    ptrs = [None] * (1 << k)
                    ptrs[lam] = self._peers[rank_of[lam]][0]
                    self.shard.swap_bits_p2p(lbits, beta, ptrs, 2 * self.EXCHANGE_SMS)
            else:
                self._swap_staged(lbits, beta, {lam: rank_of[lam]})
        if device:
            with torch.cuda.stream(comm):
                self._stream_barrier()      # every transfer has landed everywhere
                done = torch.cuda.Event()
                done.record(comm)
            cur.wait_event(done)
        self._xch_end(t_begin)
        self.shard.passes_launched += len(passes)
        self.shard.gates_applied += sum(len(p.ops) + len(p.tail) for p in passes)
        for pg, pl in pairs:
            qa, qb = self._logical_at(pg), self._logical_at(pl)
            self.pos[qa], self.pos[qb] = pl, pg
        self.swaps += k
        self.overlapped_exchanges += 1
        if k > 1:
            self.fused_exchanges += 1
        self.swap_bytes += 16 * (1 << (self.n_local - k)) * ((1 << k) - 1)

    def _apply_local(self, run):
        """Apply localised ops (GateOps on local bits and ("scale", factor) items: diagonal gates whose qubits are all
        global).  A scalar commutes with everything, so the factors of a batch are multiplied up and folded into the
        matrix of one of its gates -- no launch, and the batch is not cut into pieces at every scalar (a QFT's
        controlled phases between global qubits used to split its Hadamard passes into single-gate passes)."""
        batch, factor = [], 1.0 + 0.0j
        for item in run:
            if isinstance(item, tuple):
                factor *= complex(item[1])
            else:
                batch.append(item)
        if factor != 1.0:
            for i, op in enumerate(batch):
                if op.kind in (_lib.GATE_DENSE1, _lib.GATE_DENSE2, _lib.GATE_DIAG1, _lib.GATE_DIAG2):
                    batch[i] = GateOp(op.kind, op.qubits, np.asarray(op.matrix, dtype=complex) * factor)
                    break
            else:
                self.shard.scale(factor)
        if batch:
            self.shard.apply_ops(batch)

    def _swap_in_for(self, deferred):
        """Bring the global qubits the deferred gates need into local bits."""
        pairs = self._plan_swap(deferred)
        if pairs:
            self.swap_many(pairs)

    def _plan_swap(self, deferred):
        """[(physical global bit, physical local victim bit), ...] for the global qubits the deferred gates need."""
        # a deferred SWAP marker renames its two qubits for everything behind it: look at the deferred gates through the
        # names their qubits have NOW
        alias, seen_by = {}, []
        for op in deferred:
            if op.kind == "swap":
                a, b = op.qubits
                alias[a], alias[b] = alias.get(b, b), alias.get(a, a)
                seen_by.append(None)
            else:
                seen_by.append((op, [alias.get(q, q) for q in op.qubits]))
        first_use, order = {}, []
        for idx, item in enumerate(seen_by):
            for q in (item[1] if item else ()):
                if q not in first_use:
                    first_use[q] = idx
                    order.append(q)
        need = []
        for item in seen_by:
            if item is None:
                continue
            op, qubits = item
            for k, q in enumerate(qubits):
                if self.pos[q] < self.n_local or q in need:
                    continue
                diag_ok = op.kind in (_lib.GATE_DIAG1, _lib.GATE_DIAG2) or (op.kind == _lib.GATE_CX and k == 0)
                if not diag_ok:
                    need.append(q)
        need.sort(key=lambda q: first_use[q])
        never = len(deferred) + 1
        pairs = []
        taken = set()
        for q in need[: self.g]:
            if self.pos[q] < self.n_local:
                continue
            # victim: local logical qubit used furthest in the future (and not a partner of q's next gate)
            partners = set()
            for item in seen_by:
                if item and q in item[1]:
                    partners = set(item[1])
                    break
            best, best_use = None, -1
            for lq in range(self.n):
                if self.pos[lq] >= self.n_local or lq in need or lq in partners or lq in taken:
                    continue
                # the pinned low bits stay local when passes are to overlap the exchange: they are tile bits of every
                # pass and a sub-block pass cannot have a swapped bit in its tile
                # (and exchanging one of them with a peer would move pieces of 16 ... 128 bytes over NVLink)
                if self.pos[lq] < _LOW_BITS and self.n_local - _LOW_BITS > 2 * self.g:
                    continue
                use = first_use.get(lq, never)
                if use > best_use:
                    best, best_use = lq, use
            if best is None:
                if pairs:  # tiny shards: this one comes in with a later exchange, after the gates of the others ran
                    continue
                raise RuntimeError("no local qubit available to swap out")
            taken.add(best)
            pairs.append((self.pos[q], self.pos[best]))
        return pairs

    def ensure_local(self, qubit):
        """Make logical ``qubit`` a local bit (used before measure / reset on it)."""
        if self.pos[qubit] < self.n_local:
            return
        # swap with the top local bit's occupant
        self.swap_global_local(self.pos[qubit], self.n_local - 1)

    def restore_identity_layout(self):
        """Undo every remap so that logical qubit q sits at physical bit q again (needed before the
        state is read out or sampled in logical index order)."""
        # home qubits that sit in ANOTHER global bit are brought into a local bit first (one at a time) ...
        for phys in range(self.n_local, self.n):
            q_home = phys  # logical qubit whose home is this global bit
            if self.pos[q_home] != phys and self.pos[q_home] >= self.n_local:
                homes = {self.pos[q] for q in range(self.n_local, self.n)}
                victim = max(p for p in range(self.n_local) if p not in homes)
                self.swap_global_local(self.pos[q_home], victim)
        # ... then every misplaced global bit receives its home qubit in ONE fused exchange
        swap = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=complex)
        pairs = [(phys, self.pos[phys]) for phys in range(self.n_local, self.n) if self.pos[phys] != phys]
        # a home qubit that sits in one of the lowest local bits (a QFT's final swaps send qubit n-1 to bit 0) first moves
        # to a high local bit inside the shard: exchanging bit 0 with a peer would move 16-byte pieces over NVLink
        if any(pl < _LOW_BITS for _, pl in pairs) and self.n_local - _LOW_BITS > 2 * self.g:
            busy = {pl for _, pl in pairs}
            free_high = [p for p in range(self.n_local - 1, _LOW_BITS - 1, -1) if p not in busy]
            ops = []
            for (_, pl), high in zip([pr for pr in pairs if pr[1] < _LOW_BITS], free_high):
                ops.append(GateOp(_lib.GATE_DENSE2, [pl, high], swap))
                qa, qb = self._logical_at(pl), self._logical_at(high)
                self.pos[qa], self.pos[qb] = high, pl
            self.shard.apply_ops(ops)
            pairs = [(phys, self.pos[phys]) for phys in range(self.n_local, self.n) if self.pos[phys] != phys]
        if pairs:
            self.swap_many(pairs)
        # now fix the local permutation with SWAP gates
        ops = []
        pos = list(self.pos)
        for q in range(self.n_local):
            if pos[q] == q:
                continue
            other = pos.index(q)  # logical qubit currently sitting at physical bit q
            ops.append(GateOp(_lib.GATE_DENSE2, [pos[q], q], swap))
            pos[other], pos[q] = pos[q], q
        if ops:
            self.shard.apply_ops(ops)
        self.pos = pos
        assert self.pos == list(range(self.n))

    # -- measurement -------------------------------------------------------------------
    def _allreduce(self, values):
        torch = self._torch
        t = torch.tensor(values, dtype=torch.float64, device=self.shard.comm_device())
        self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM, group=self.group)
        return t.cpu().tolist()

    def prob_qubit(self, qubit):
        phys = self.pos[qubit]
        if phys < self.n_local:
            p0, p1 = self.shard.prob_qubit(phys)
        else:
            nrm = self.shard.norm2()
            p0, p1 = (0.0, nrm) if self._rank_bit(phys) else (nrm, 0.0)
        return tuple(self._allreduce([p0, p1]))

    def norm2(self):
        return self._allreduce([self.shard.norm2()])[0]

    def expval_pauli(self, pauli):
        """<psi|P|psi> for a Pauli label ("XIZY": leftmost character = highest LOGICAL qubit), as
        engine.DeviceState.expval_pauli (Statevector._expectation_value_pauli, statevector.py:377-408).  Qubits
        with X or Y are made local first (victims: local qubits with I or Z); a Z on a global qubit is the sign
        of the rank's bit; the ranks' partial sums are all-reduced."""
        label = pauli.upper()
        if len(label) != self.n or any(c not in "IXYZ" for c in label):
            raise ValueError("Pauli label must have %d characters from IXYZ" % self.n)
        char = {q: c for q, c in enumerate(reversed(label))}
        flips = [q for q in range(self.n) if char[q] in "XY"]
        if len(flips) > self.n_local:
            raise _lib.QsvLibraryError("a Pauli with %d X/Y factors does not fit in a shard of %d qubits"
                                       % (len(flips), self.n_local))
        for q in flips:
            if self.pos[q] >= self.n_local:
                victim = max(p for p in range(self.n_local) if char[self._logical_at(p)] in "IZ")
                self.swap_global_local(self.pos[q], victim)
        if all(c == "I" for c in label):
            return float(np.sqrt(self.norm2()))  # the reference returns the norm here (statevector.py:397-398)
        local = "".join(char[self._logical_at(p)] for p in reversed(range(self.n_local)))
        sign = 1.0
        for p in range(self.n_local, self.n):
            if char[self._logical_at(p)] == "Z" and self._rank_bit(p):
                sign = -sign
        part = self.shard.norm2() if all(c == "I" for c in local) else self.shard.expval_pauli(local)
        return self._allreduce([sign * part])[0]

    def collapse(self, qubit, outcome, prob, is_reset=False):
        self.ensure_local(qubit)
        self.shard.collapse(self.pos[qubit], outcome, prob, is_reset)

    def sample_all(self, uniforms, check_norm=True):
        """All n qubits, in the reference's bin order and with ITS running sums: numpy's ``choice`` is
        ``cdf = p.cumsum(); cdf /= cdf[-1]; searchsorted(u, 'right')`` (qasm_simulator.py:213), one sequential
        float64 sum over the 2^n bins.  With the identity layout the bins of rank r follow those of rank r - 1, so
        the sum is CHAINED through the ranks: every rank classifies the chunks of its shard in parallel
        (qsv_cdf_prepare / qsv_cdf_classify, steered by the all-gathered approximate totals), then rank r continues
        the exact running sum where rank r - 1 stopped (qsv_cdf_chain: milliseconds, one double handed on), and
        finally every rank looks the uniforms up in its own share of the CDF (qsv_cdf_search).  Seeded counts are
        the reference's bit for bit, as on one GPU.  Returns (logical bin indices, total) on every rank."""
        torch = self._torch
        dist = self._dist
        self.restore_identity_layout()
        uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
        shots = int(uniforms.size)
        dev = self.shard.comm_device()
        approx = torch.zeros(self.world, dtype=torch.float64, device=dev)
        approx[self.rank] = self.shard.cdf_prepare(shots)
        dist.all_reduce(approx, op=dist.ReduceOp.SUM, group=self.group)
        approx = approx.cpu().tolist()
        self.shard.cdf_classify(float(sum(approx[: self.rank])), shots)
        start, end = 0.0, 0.0
        for r in range(self.world):  # the chain: rank r starts from the exact end of rank r - 1
            handoff = torch.zeros(1, dtype=torch.float64, device=dev)
            if r == self.rank:
                start = end
                handoff[0] = self.shard.cdf_chain(start, shots)
            dist.broadcast(handoff, src=self._global_rank(r), group=self.group)
            end = float(handoff.item())
        total = end
        if check_norm and not abs(total - 1.0) <= 1.4901161193847656e-08:
            raise ValueError("probabilities do not sum to 1")  # numpy's choice()
        if not total > 0.0:
            raise ValueError("probabilities sum to zero")
        idx = self.shard.cdf_search(start, total, uniforms)
        mine = idx >= 0
        phys = np.where(mine, idx + (self.rank << self.n_local), 0).astype(np.int64)
        out = torch.as_tensor(np.stack([phys, mine.astype(np.int64)]), device=dev)
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=self.group)
        out = out.cpu().numpy()
        if not np.all(out[1] == 1):  # every shot lies in exactly one shard's share of the CDF
            raise _lib.QsvLibraryError("sharded sampler: a shot was claimed by %d shards" % int(out[1].max()))
        return out[0], total

    # -- the interface simulator.py drives (same method set as engine.DeviceState), so that the BasicAer
    # -- mirror classes run SPMD on a sharded state: every rank executes the same run(qobj) ----------
    MAX_MARGINAL_QUBITS = 28  # 2 GiB of float64 bins per rank

    @property
    def passes_launched(self):
        return getattr(self.shard, "passes_launched", None)

    @property
    def gates_applied(self):
        return getattr(self.shard, "gates_applied", None)

    def upload(self, amplitudes, phase=1.0 + 0.0j):
        """Every rank holds the full host vector (it comes from the Qobj) and uploads its own slab."""
        amps = np.asarray(amplitudes, dtype=complex).reshape(-1)
        if amps.size != (1 << self.n):
            raise _lib.QsvLibraryError("upload: expected 2^%d amplitudes" % self.n)
        self.pos = list(range(self.n))
        size = 1 << self.n_local
        self.shard.upload(amps[self.rank * size:(self.rank + 1) * size], phase)

    def marginal(self, measured_sorted):
        """Marginal probabilities of the measured LOGICAL qubits (bin bit pos <-> measured_sorted[pos], the
        reference's layout, qasm_simulator.py:202-209), identical on every rank: the rank's own marginal over
        its local measured bits is scattered to the bins its rank bits select, then all-reduced."""
        torch = self._torch
        m = len(measured_sorted)
        if m > self.MAX_MARGINAL_QUBITS:
            raise _lib.QsvLibraryError("marginal over %d qubits does not fit on a rank" % m)
        dev = self.shard.comm_device()
        local = [(i, self.pos[q]) for i, q in enumerate(measured_sorted) if self.pos[q] < self.n_local]
        glob = [(i, self.pos[q]) for i, q in enumerate(measured_sorted) if self.pos[q] >= self.n_local]
        phys_sorted = sorted(p for _, p in local)
        probs_local = self.shard.marginal(phys_sorted)
        dest_of = {p: i for i, p in local}
        base = sum(self._rank_bit(p) << i for i, p in glob)
        # bit j of the local bin index (physical bit phys_sorted[j]) goes to bit dest_of[...] of the full index;
        # the rank's global measured bits select the block (qsv_scatter_bits)
        full = self.shard.scatter_bits(probs_local, [dest_of[p] for p in phys_sorted], base, m).to(dev)
        self._dist.all_reduce(full, op=self._dist.ReduceOp.SUM, group=self.group)
        return full

    def sample(self, measured_sorted, uniforms, exact_order=True, check_norm=True):
        """engine.DeviceState.sample on the sharded state; every rank returns the same bin indices.
        Fewer than all qubits measured: the all-reduced marginal is sampled with the single-GPU sampler
        (numpy's cumsum order when ``exact_order``), on every rank alike.  All qubits measured (or a marginal
        too large for one rank): ``sample_all`` -- the exact-order CDF chained through the ranks."""
        measured_sorted = list(measured_sorted)
        m = len(measured_sorted)
        if m == self.n or m > self.MAX_MARGINAL_QUBITS:
            logical, total = self.sample_all(uniforms, check_norm=check_norm)
            if m == self.n:
                return logical, total
            bins = np.zeros_like(logical)
            for i, q in enumerate(measured_sorted):
                bins |= ((logical >> q) & 1) << i
            return bins, total
        probs = self.marginal(measured_sorted)
        samples, total = self.shard.sample_probs(probs, m, uniforms, exact_order=exact_order, check_norm=check_norm)
        # every rank computed the same thing from the same all-reduced marginal; rank 0's copy is authoritative
        torch = self._torch
        t = torch.as_tensor(np.ascontiguousarray(samples), device=self.shard.comm_device())
        self._dist.broadcast(t, src=self._dist.get_global_rank(self.group, 0) if self.group is not None else 0,
                             group=self.group)
        return t.cpu().numpy(), total

    def chop(self, threshold):
        self.shard.chop(threshold)

    MAX_GATHER_QUBITS = 31  # 32 GiB of complex128 per rank on the host

    def download(self):
        """Full statevector in logical order on every rank (sizes that fit in host memory)."""
        if self.n > self.MAX_GATHER_QUBITS:
            raise _lib.QsvLibraryError(
                "the full statevector of %d qubits (%d GiB) is not gathered on every rank; sample it "
                "(qasm_simulator) or read the shards" % (self.n, 16 * 2 ** self.n >> 30))
        return self.gather_statevector()

    def snapshot(self):
        return (self.shard.snapshot(), list(self.pos))

    def restore(self, snap):
        self.shard.restore(snap[0])
        self.pos = list(snap[1])

    def _apply_dense_k(self, op):
        """k >= 3 qubit ``unitary``: make its qubits local (victims: the highest local bits it does not use),
        then the shard's dense k-qubit kernel."""
        for q in op.qubits:
            if self.pos[q] >= self.n_local:
                used = {self.pos[x] for x in op.qubits}
                victim = max(p for p in range(self.n_local) if p not in used)
                self.swap_global_local(self.pos[q], victim)  # updates self.pos
        self.shard.apply_ops([GateOp(op.kind, [self.pos[q] for q in op.qubits], op.matrix)])

    def gather_statevector(self):
        """Full logical-order statevector on every rank (small n only; tests)."""
        torch = self._torch
        self.restore_identity_layout()
        local = torch.as_tensor(self.shard.download().view(np.float64), device=self.shard.comm_device())
        parts = [torch.empty_like(local) for _ in range(self.world)]
        self._dist.all_gather(parts, local, group=self.group)
        return np.concatenate([p.cpu().numpy().view(np.complex128) for p in parts])


# ------------------------------------------------------------------------------------------------
def distributed_simulator(backend="qasm_simulator", shard_factory=None, group=None,
                          chunk_amps=DEFAULT_CHUNK_AMPS, **kwargs):
    """The BasicAer mirror simulators (simulator.QasmSimulatorB200 / StatevectorSimulatorB200) on a state
    sharded over the ranks of ``group`` (default: the world).  SPMD: every rank constructs the simulator and
    calls ``run(qobj)`` with the same Qobj; every rank returns the same result dict (counts, memory and -- for
    sizes that fit in host memory -- the statevector).  The qubit cap grows by log2(world) over the
    single-GPU cap.  ``torch.distributed`` must be initialised (NCCL on GPUs); ``shard_factory(n_local)``
    defaults to engine.DeviceState on the current CUDA device.  A seed the backend draws itself is rank 0's."""
    import torch
    import torch.distributed as dist

    from . import simulator
    from .engine import DeviceState

    world = dist.get_world_size(group)
    g = int(round(math.log2(world)))
    if backend not in ("qasm_simulator", "statevector_simulator", "unitary_simulator"):
        raise LookupError("The '%s' backend is not installed in your system." % backend)
    factory = shard_factory or DeviceState
    if backend == "unitary_simulator":  # a 2n-qubit vector, sharded the same way
        def unitary_state_factory(n2):
            if n2 - g < 2:
                raise simulator.BasicAerError("a unitary on a %d-rank sharded state needs at least %d qubits"
                                              % (world, (g + 3) // 2))
            return ShardedState(n2, factory, chunk_amps=chunk_amps, group=group)

        kwargs.setdefault("max_qubits", (simulator.max_qubits_for_device() + g) // 2)
        return simulator.UnitarySimulatorB200(state_factory=unitary_state_factory, **kwargs)

    def state_factory(n_qubits):
        if n_qubits - g < 2:  # tiny circuits: pad with idle qubits so that every rank holds >= 2 local qubits
            raise simulator.BasicAerError("a circuit on a %d-rank sharded state needs at least %d qubits"
                                          % (world, g + 2))
        return ShardedState(n_qubits, factory, chunk_amps=chunk_amps, group=group)

    def seed_sync(seed):
        dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
        t = torch.tensor([seed], dtype=torch.int64, device=dev)
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        return int(t.item())

    cls = simulator.QasmSimulatorB200 if backend == "qasm_simulator" else simulator.StatevectorSimulatorB200
    kwargs.setdefault("max_qubits", simulator.max_qubits_for_device() + g)
    return cls(state_factory=state_factory, seed_sync=seed_sync, **kwargs)


# Golden fixtures (outputs of the unmodified reference, tests/golden/) run through the sharded simulators before the
# timed region of bench.py --gpus N: seeded counts / memory identical, amplitudes <= 1e-10.  A case is used when its
# register leaves every rank at least two local qubits.
PARITY_CASES = ["qv_20_counts", "qv_24_counts", "qft_20", "qv_18_state", "midcircuit_8", "sv_midcircuit_10",
                "transpiled_init_reset_cond_counts", "sampler_marginal_subset", "random_14_counts", "random_6_unitary",
                "random_4_unitary", "unitary_random"]


def parity_gate(factory_kwargs, world, cases=None, device="cuda"):
    """Run PARITY_CASES through distributed_simulator on every rank; returns (summary string, details)."""
    import sys
    import time

    import torch
    import torch.distributed as dist

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tests_dir = os.path.join(root, "tests")
    if tests_dir not in sys.path:
        sys.path.insert(0, tests_dir)
    import golden_io  # fixtures + comparison helpers are test infrastructure (no oracle code involved)
    import helpers

    g = int(round(math.log2(world)))
    ran, failed, details = 0, [], {}
    for name in (cases or PARITY_CASES):
        case = golden_io.load_case(name)
        n = case["qobj"]["config"]["n_qubits"]
        n_state = 2 * n if case["backend"] == "unitary_simulator" else n
        if n_state - g < 2:
            continue
        t0 = time.perf_counter()
        err = ""
        try:
            result = helpers.run_case_with(lambda backend: distributed_simulator(backend, **factory_kwargs), case)
            helpers.compare_result(result, case, amp_tol=1e-10)
        except AssertionError as exc:
            err = "mismatch %s" % (str(exc)[:120],)
        flag = torch.tensor([1 if err else 0], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)  # a mismatch on ANY rank fails the case
        ran += 1
        details[name] = {"backend": case["backend"], "n_qubits": n, "ok": int(flag.item()) == 0,
                         "s": round(time.perf_counter() - t0, 3)}
        if int(flag.item()):
            failed.append(name + (": " + err if err else ""))
    summary = "golden %d/%d identical" % (ran - len(failed), ran)
    if failed:
        summary += " FAILED: " + "; ".join(failed)
    return summary, details


def bench_main(args, bench):
    """bench.py leg for N > 1 GPUs: weak scaling, args.qubits local qubits per GPU.  ``bench`` is the bench.py module
    (metric names, clock sampler, peaks, workload description)."""
    import json

    import torch
    import torch.distributed as dist

    from . import circuits, lower
    from .engine import DeviceState

    world = int(os.environ["WORLD_SIZE"])
    rank = int(os.environ["RANK"])
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    g = int(round(math.log2(world)))
    free, _ = torch.cuda.mem_get_info()
    n_local = args.qubits
    chunk = DEFAULT_CHUNK_AMPS
    while n_local > 20 and (16 << n_local) + 4 * 16 * min(chunk, 1 << (n_local - 1)) + (2 << 30) > free:
        n_local -= 1
    nl = torch.tensor([n_local], device="cuda")
    dist.all_reduce(nl, op=dist.ReduceOp.MIN)
    n_local = int(nl.item())
    n = n_local + g
    depth = args.depth or n
    tile = args.tile or None
    shots, qv_seed, sim_seed = bench.SHOTS, bench.QV_SEED, bench.SIM_SEED
    factory = (lambda k: DeviceState(k, tile_qubits=tile)) if tile else DeviceState
    chunk_amps = min(chunk, 1 << (n_local - 1))

    # ---- parity gate: golden fixtures through the sharded simulators (every rank, before anything is timed) ----
    parity, parity_details = ("skipped (--no-parity)", {})
    if not args.no_parity:
        parity, parity_details = parity_gate({"shard_factory": factory, "chunk_amps": 1 << 22, "max_qubits": 40}, world)
        torch.cuda.empty_cache()

    ins = [i for i in circuits.quantum_volume_instructions(n, depth, qv_seed) if i["name"] == "unitary"]
    ops = [lower.lower_gate(i["name"], i["qubits"], i.get("params")) for i in ins]
    n_gates = len(ops)
    overlap = os.environ.get("QSV_DIST_OVERLAP", "1") == "1"  # A/B knob of tools/: exchanges without late passes
    st = ShardedState(n, factory, chunk_amps=chunk_amps, overlap=overlap)
    if os.environ.get("QSV_DIST_OVERLAP_GATES"):
        st.OVERLAP_TARGET_GATES = int(os.environ["QSV_DIST_OVERLAP_GATES"])
    if os.environ.get("QSV_DIST_CARRY"):      # A/B knobs of tools/: carry-over of the gates an exchange does not need,
        st.carry_over = os.environ["QSV_DIST_CARRY"] == "1"
    if os.environ.get("QSV_DIST_MIN_OVL_SWAPS"):  # 1 = pairwise swaps overlap their late passes too
        st.MIN_OVERLAP_SWAPS = int(os.environ["QSV_DIST_MIN_OVL_SWAPS"])
    if os.environ.get("QSV_DIST_MIN_OVL_GATES"):
        st.MIN_OVERLAP_GATES = int(os.environ["QSV_DIST_MIN_OVL_GATES"])
    if os.environ.get("QSV_DIST_STEPWISE"):
        st.p2p_stepwise = os.environ["QSV_DIST_STEPWISE"] == "1"
    if os.environ.get("QSV_DIST_XSMS"):       # SMs left to the exchange kernel under the sub-block passes
        st.EXCHANGE_SMS = int(os.environ["QSV_DIST_XSMS"])
    uniforms = np.random.RandomState(sim_seed).random_sample(shots)

    def step():
        st.init_basis0()
        st.apply_ops(ops)
        return st.sample_all(uniforms)  # restores the identity layout, then the exact-order CDF across the ranks

    def timed(fn, reps):
        """Device time of `reps` calls of fn, max over ranks (CUDA events, barrier + synchronize on both sides)."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        dist.barrier()
        e0.record()
        out = None
        for _ in range(reps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / reps, out

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    clocks = bench.ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = _lib.launch_count()
    swaps0, bytes0, fused0 = st.swaps, st.swap_bytes, st.fused_exchanges
    passes0 = st.shard.passes_launched
    xch0 = st.exchange_ms
    ovl0 = st.overlapped_exchanges
    wait0 = st.exchange_wait_ms()
    step_ms, (samples, total) = timed(step, args.steps)
    launches = torch.tensor([_lib.launch_count() - launches0], device="cuda", dtype=torch.int64)
    dist.all_reduce(launches, op=dist.ReduceOp.SUM)
    passes_dev = (st.shard.passes_launched - passes0) / args.steps
    swaps_dev = (st.swaps - swaps0) / args.steps
    fused_dev = (st.fused_exchanges - fused0) / args.steps
    swap_gb_dev = (st.swap_bytes - bytes0) / args.steps / 1e9
    exchange_ms = st.exchange_time_ms(xch0) / args.steps
    st_overlapped = st.overlapped_exchanges - ovl0
    per_rank = torch.tensor([exchange_ms, (st.exchange_wait_ms() - wait0) / args.steps], device="cuda", dtype=torch.float64)
    per_rank_all = [torch.zeros_like(per_rank) for _ in range(world)]
    dist.all_gather(per_rank_all, per_rank)
    per_rank_all = [[round(float(v), 1) for v in t.tolist()] for t in per_rank_all]
    clock_info = clocks.stop() if rank == 0 else None
    st_carry, st_xsms, st_stepwise = st.carry_over, st.EXCHANGE_SMS, st.p2p_stepwise

    # ---- the other circuit families of BASELINE's metric on the sharded state (one warm-up, one timed run each) ----
    families = {}
    if not args.no_families:
        qft_ins = [i for i in circuits.qft(n, state_seed=3)["experiments"][0]["instructions"]
                   if i["name"] not in ("measure", "barrier")]
        rnd_ins = circuits.random_basis_instructions(n, bench.RANDOM_DEPTH, 7)
        for key, label, fins in (("qft", "QFT(%d) with final swaps on a product state, basis gates" % n, qft_ins),
                                 ("random_circuit", "random basis-gate circuit, %d qubits, depth %d, seed 7"
                                  % (n, bench.RANDOM_DEPTH), rnd_ins)):
            fops = [o for o in (lower.lower_gate(i["name"], i["qubits"], i.get("params")) for i in fins) if o is not None]

            def run_family():
                st.init_basis0()
                st.apply_ops(fops)
                st.restore_identity_layout()
                return st.norm2()

            run_family()
            p0, s0, b0 = st.shard.passes_launched, st.swaps, st.swap_bytes
            fam_ms, nrm = timed(run_family, 1)
            families[key] = {"workload": label, "instructions": len(fops), "circuit_s": fam_ms * 1e-3,
                             "local_passes": st.shard.passes_launched - p0, "qubit_swaps": st.swaps - s0,
                             "swap_GB_per_gpu": (st.swap_bytes - b0) / 1e9, "norm2": nrm,
                             "reference_equivalent_GBps": len(fops) * 32.0 * 2 ** n / (fam_ms * 1e-3) / 1e9}

    # ---- end-to-end leg: the same circuit through the plugin-facing call on every rank (SPMD), host Qobj in
    # (gate matrices + shots), counts out; wall clock around run(), max over ranks -----------------------------
    e2e, e2e_error = None, None
    if not args.no_e2e and os.environ.get("QSV_BENCH_NO_DIST_E2E") != "1":
        import copy
        import time

        st.close()  # unmap the peers' shards before anybody frees one
        del st
        torch.cuda.empty_cache()
        try:
            qobj = circuits.quantum_volume(n, depth, qv_seed, shots=shots, seed_simulator=sim_seed)
            sim = distributed_simulator("qasm_simulator", shard_factory=factory, chunk_amps=chunk_amps, max_qubits=n)
            sim.run(copy.deepcopy(qobj))  # warm-up (allocation of the shard and the staging buffers)
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                res = sim.run(copy.deepcopy(qobj))
            torch.cuda.synchronize()
            dist.barrier()
            e2e_ms = torch.tensor([(time.perf_counter() - t0) * 1e3 / args.steps], device="cuda", dtype=torch.float64)
            dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
            counts = res["results"][0]["data"]["counts"]
            assert sum(counts.values()) == shots
            e2e_s = float(e2e_ms.item()) * 1e-3
            e2e = {"value": n_gates * 32.0 * 2 ** n / e2e_s / 1e9, "unit": bench.UNIT,
                   "h2d_bytes_per_step": int(world * (n_gates * 256 + shots * 8)),
                   "d2h_bytes_per_step": int(world * (shots * 8 + 8)), "circuit_s": e2e_s,
                   "api": "dist.distributed_simulator('qasm_simulator').run(qobj dict) -> counts, on every rank "
                          "(SPMD); the final layout is restored before sampling so that counts follow the "
                          "reference's bin order"}
        except (_lib.QsvLibraryError, RuntimeError, TypeError, ValueError, KeyError, AttributeError,
                AssertionError) as exc:  # RuntimeError: e.g. an out-of-memory error, alike on all ranks
            # a deterministic host-side failure is the same on every rank: report it instead of losing the line
            e2e_error = "%s: %s" % (type(exc).__name__, exc)
    if rank == 0:
        bytes_per_gate_pass = 32.0 * 2 ** n
        value = n_gates * bytes_per_gate_pass / (step_ms * 1e-3) / 1e9
        peak, peak_src = bench.measured_peaks()
        cfg = bench.workload_config(n, depth, args.tile or DEFAULT_TILE, world, n_local)
        line = {
            "metric": bench.METRIC, "value": value, "unit": bench.UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "circuit_s": step_ms * 1e-3,
            "schedule": {"local_passes_per_step": passes_dev, "qubit_swaps_per_step": swaps_dev,
                         "fused_exchanges_per_step": fused_dev, "swap_GB_per_gpu_per_step": swap_gb_dev,
                         "overlapped_exchanges_per_step": (st_overlapped / args.steps),
                         "exchange_ms_per_step": exchange_ms,
                         "exchange_ms_per_step_by_rank": [v[0] for v in per_rank_all],
                         "exchange_barrier_wait_ms_per_step_by_rank": [v[1] for v in per_rank_all],
                         "carry_over": bool(st_carry), "exchange_sms": st_xsms, "p2p_stepwise": bool(st_stepwise),
                         "exchange_GBps_per_gpu_per_direction": (swap_gb_dev / (exchange_ms * 1e-3) if exchange_ms else None),
                         "nvlink_peak_GBps_per_direction": 900.0},
            "roofline": {"bound": "hbm", "kernel": "tile_pass_kernel", "achieved": None, "peak": peak, "unit": "GB/s",
                         "frac": None, "traffic": None, "peak_source": peak_src,
                         "note": "per-kernel roofline is reported by the N=1 run"},
            "cpu_baseline": None,
            "e2e": e2e or {"value": value, "unit": bench.UNIT, "h2d_bytes_per_step": int(n_gates * 256 + shots * 8),
                           "d2h_bytes_per_step": int(shots * 8 + 8 * world),
                           "note": "e2e leg %s: the device-timed leg is driven from host op lists; uniforms in, "
                                   "sample indices out are inside its timed region"
                                   % ("failed (%s)" % e2e_error if e2e_error else "skipped")},
            "qft": families.get("qft"),
            "random_circuit": families.get("random_circuit"),
            "gpu_launches": int(launches.item()),
            "clocks": clock_info,
            "check": {"sample_total_prob": total, "parity": parity, "parity_cases": parity_details,
                      "sampler": "exact-order CDF chained through the ranks (counts bit-identical to the reference's "
                                 "seeded RandomState.choice)"},
        }
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()
