"""Multi-rank path on CPU: world_size 2 and 4 over gloo, the shard replaced by the numpy test double.
Exercises the sharding logic of dist.ShardedState -- localisation of gates on global qubits, qubit
swaps (pack / exchange / unpack), layout restore, distributed P(qubit), collapse and sampling --
against the single-process oracle."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, seed, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fake_state import FakeState
    from oracle import basicaer_oracle as orc
    from qiskit_sdk_py_b200 import _lib, circuits, lower
    from qiskit_sdk_py_b200.dist import ShardedState

    ins = circuits.random_basis_instructions(n, 10, seed) + circuits.quantum_volume_instructions(n, 3, seed)
    ops = [lower.lower_gate(i["name"], i["qubits"], i.get("params")) for i in ins]
    ops = [o for o in ops if o is not None]
    st = ShardedState(n, lambda k: FakeState(k, tile_qubits=4), chunk_amps=8)
    st.init_basis0(np.exp(0.2j))
    st.apply_ops(ops)
    # oracle
    ref = np.zeros(2 ** n, dtype=complex)
    ref[0] = np.exp(0.2j)
    for o in ops:
        if o.kind in (_lib.GATE_DIAG1, _lib.GATE_DIAG2):
            mat = np.diag(o.matrix)
        elif o.kind == _lib.GATE_CX:
            mat = orc.cx_gate_matrix()
        else:
            mat = o.matrix
        ref = orc.apply_unitary(ref.reshape([2] * n), mat, list(o.qubits)).reshape(-1)
    checks = {"swaps": st.swaps, "fused": st.fused_exchanges}
    # distributed probabilities on a permuted layout
    p = np.abs(ref) ** 2
    idx = np.arange(2 ** n)
    for q in range(n):
        p0, p1 = st.prob_qubit(q)
        mask = ((idx >> q) & 1).astype(bool)
        assert abs(p0 - p[~mask].sum()) < 1e-12 and abs(p1 - p[mask].sum()) < 1e-12, (q, p0, p1)
    assert abs(st.norm2() - 1.0) < 1e-12
    # sampling on the permuted layout: indices must be distributed like p (logical order)
    u = np.random.RandomState(7).random_sample(4000)
    samples, total = st.sample_all(u)
    assert abs(total - 1.0) < 1e-12
    hist = np.bincount(samples, minlength=2 ** n) / 4000.0
    assert np.abs(hist - p).max() < 0.05
    # collapse a (possibly global) qubit
    qm = n - 1
    probs = st.prob_qubit(qm)
    st.collapse(qm, 1, probs[1], is_reset=False)
    mask = ((idx >> qm) & 1).astype(bool)
    ref2 = np.where(mask, ref, 0) / np.sqrt(probs[1])
    full = st.gather_statevector()
    assert st.pos == list(range(n))
    err = float(np.max(np.abs(full - ref2)))
    assert err < 1e-12, err
    checks["err"] = err
    if rank == 0:
        np.save(os.path.join(out_dir, "ok_%d_%d.npy" % (world, n)), np.array([checks["swaps"], err, checks["fused"]]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 5), (2, 7), (4, 6), (4, 8)])
def test_sharded_state_matches_oracle(tmp_path, world, n):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, 11 + n, str(tmp_path)), nprocs=world, join=True)
    res = np.load(os.path.join(str(tmp_path), "ok_%d_%d.npy" % (world, n)))
    assert res[0] > 0, "the test circuit should have needed qubit swaps"
    assert res[1] < 1e-12
    if world == 4:
        assert res[2] > 0, "world 4 should have used the fused multi-qubit exchange"
