#!/usr/bin/env python
"""Dry run of the sharded schedule (no GPU, no torch.distributed): ShardedState.apply_ops on rank 0 of a
virtual world with a counting shard -- local passes (and their kernel ops after gate fusion, from the host-only
qsv_pass_info), exchanges and exchanged bytes -- plus a time estimate from the measured single-GPU numbers.
    python tools/dist_dry_run.py [world] [local_qubits]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from qiskit_sdk_py_b200 import circuits, lower
from qiskit_sdk_py_b200.dist import ShardedState
from qiskit_sdk_py_b200.engine import DeviceState
from qiskit_sdk_py_b200.planner import merge_ops, plan_passes, DEFAULT_TILE_QUBITS


class CountingShard:
    def __init__(self, n):
        self.n = n
        self.passes = []  # (gates, kernel ops, fused blocks)
        self.late = []    # ... of the passes that ran under an exchange

    def init_basis0(self, phase=1.0):
        pass

    def scale(self, f):
        pass

    tile_qubits = DEFAULT_TILE_QUBITS
    passes_launched = gates_applied = 0

    def apply_ops(self, ops):
        for p in plan_passes(merge_ops(ops), self.n, DEFAULT_TILE_QUBITS, tails=True):
            info = DeviceState.pass_info(self.n, p.tile_qubits, p.ops)
            self.passes.append((len(p.ops), info["ops"], info["fused3"]))

    def apply_pass(self, tile_qubits, ops, dest=None, tail=None):
        info = DeviceState.pass_info(self.n, tile_qubits, ops)
        self.passes.append((len(ops), info["ops"], info["fused3"]))

    def plan_ops(self, ops, avoid=()):
        return plan_passes(merge_ops(list(ops)), self.n, DEFAULT_TILE_QUBITS, tails=True, avoid=avoid)

    def apply_pass_sub(self, tile_qubits, ops, tail, fixed_qubits, pattern, max_sms=0):
        if pattern == 0:  # count every late pass once (it runs on all 2^k sub-blocks = once over the shard)
            info = DeviceState.pass_info(self.n, tile_qubits, ops)
            self.passes.append((len(ops), info["ops"], info["fused3"]))
            self.late.append((len(ops), info["ops"], info["fused3"]))


class DryState(ShardedState):
    def __init__(self, n, world):
        import math
        self._torch = self._dist = None
        self.group = None
        self.rank, self.world = 0, world
        self.g = int(round(math.log2(world)))
        self.n, self.n_local = n, n - self.g
        self.shard = CountingShard(self.n_local)
        self.pos = list(range(n))
        self.swaps = self.fused_exchanges = self.swap_bytes = 0
        self.events = []
        self.overlap = os.environ.get("QSV_DRY_OVERLAP", "1") == "1"
        self.carry_over = os.environ.get("QSV_DRY_CARRY", "1") == "1"
        self.overlapped_exchanges = 0
        self.relabelled_swaps = 0
        self._peers = self._comm_stream = None
        self._xch_events = []
        self.late_per_exchange = []

    def _swap_staged(self, lbits, beta, peers):
        pass

    def _exchange_overlapped(self, pairs, late):
        n0 = len(self.shard.late)
        super()._exchange_overlapped(pairs, late)
        self.late_per_exchange.append((len(pairs), len(late), len(self.shard.late) - n0))
        self.events.append(("xo", len(pairs), len(self.shard.passes)))

    def swap_global_local(self, pg, pl):
        self.swap_many([(pg, pl)], force=True)

    def swap_many(self, pairs, force=False):
        k = len(pairs)
        for pg, pl in pairs:
            qa, qb = self._logical_at(pg), self._logical_at(pl)
            self.pos[qa], self.pos[qb] = pl, pg
        self.swaps += k
        self.fused_exchanges += 1
        sub = 1 << (self.n_local - k)
        self.swap_bytes += 16 * sub * ((1 << k) - 1)
        self.events.append(("x", k, len(self.shard.passes)))


def estimate(passes, swap_bytes, n_local):
    """ms: pass = max(copy, FP64) + exposed overhead, calibrated on the 33-qubit single-GPU run (68.9 ms per
    pass at 7.33 gates); exchange at 650 GB/s per GPU and direction."""
    scale = 2.0 ** (n_local - 33)
    unit_ms = 2.08 * scale      # one DMMA per 64 amplitudes over the whole shard at ~33 TF/s
    copy_ms = 43.0 * scale
    t = 0.0
    for gates, kops, blocks in passes:
        units = (kops - blocks) * 4 + blocks * 6.25
        t += max(copy_ms, units * unit_ms) + 14.0 * scale
    return t, swap_bytes / 650e9 * 1e3


if __name__ == "__main__":
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    n_local = int(sys.argv[2]) if len(sys.argv) > 2 else 33
    g = int(np.log2(world))
    n = n_local + g
    ins = [i for i in circuits.quantum_volume_instructions(n, n, 1234) if i["name"] == "unitary"]
    ops = [lower.lower_gate(i["name"], i["qubits"], i.get("params")) for i in ins]
    st = DryState(n, world)
    st.apply_ops(ops)
    ps = st.shard.passes
    t_pass, t_x = estimate(ps, st.swap_bytes, n_local)
    print("QV-%d on %d ranks (%d local): %d gates, %d passes (%.2f gates/pass), %d kernel ops (%d blocks), "
          "%d swaps in %d exchanges, %.1f GB per rank" % (n, world, n_local, len(ops), len(ps), len(ops) / len(ps),
          sum(p[1] for p in ps), sum(p[2] for p in ps), st.swaps, st.fused_exchanges, st.swap_bytes / 1e9))
    print("estimate: passes %.2f s + exchanges %.2f s = %.2f s" % (t_pass / 1e3, t_x / 1e3, (t_pass + t_x) / 1e3))
    if st.late_per_exchange:
        t_late, _ = estimate(st.shard.late, 0, n_local)
        print("overlap: %d of %d exchanges ran under late passes; (swapped qubits, late gates, late passes) per exchange: %s; "
              "late passes ~%.2f s in total" % (st.overlapped_exchanges, st.fused_exchanges + (st.swaps if world == 2 else 0),
                                               st.late_per_exchange, t_late / 1e3))
    hist = {}
    for p in ps:
        hist[p[0]] = hist.get(p[0], 0) + 1
    print("gates per pass histogram:", dict(sorted(hist.items())))
