"""torchrun worker: sharded CUDA state (NCCL) against the numpy oracle.  Launched by
tests/test_gpu_dist.py (needs >= 2 GPUs) or by hand:
    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_worker.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle import basicaer_oracle as orc  # noqa: E402
from qiskit_sdk_py_b200 import _lib, circuits, lower  # noqa: E402
from qiskit_sdk_py_b200.dist import ShardedState  # noqa: E402
from qiskit_sdk_py_b200.engine import DeviceState  # noqa: E402


def main():
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    quick = os.environ.get("QSV_DIST_WORKER_QUICK") == "1"  # by hand: only the SPMD simulator part
    for n, seed in (() if quick else ((12, 3), (16, 4), (19, 5))):
        ins = circuits.random_basis_instructions(n, 8, seed) + circuits.quantum_volume_instructions(n, 4, seed)
        ops = [lower.lower_gate(i["name"], i["qubits"], i.get("params")) for i in ins]
        ops = [o for o in ops if o is not None]
        st = ShardedState(n, DeviceState, chunk_amps=1 << 9)
        st.MIN_OVERLAP_SWAPS = 1  # exercise the overlapped exchange (sub-block passes + per-partner swaps) on 2 ranks too
        st.init_basis0(np.exp(0.3j))
        st.apply_ops(ops)
        ref = np.zeros(2 ** n, dtype=complex)
        ref[0] = np.exp(0.3j)
        for o in ops:
            if o.kind in (_lib.GATE_DIAG1, _lib.GATE_DIAG2):
                mat = np.diag(o.matrix)
            elif o.kind == _lib.GATE_CX:
                mat = orc.cx_gate_matrix()
            else:
                mat = o.matrix
            ref = orc.apply_unitary(ref.reshape([2] * n), mat, list(o.qubits)).reshape(-1)
        p = np.abs(ref) ** 2
        idx = np.arange(2 ** n)
        for q in (0, n // 2, n - 1):
            p0, p1 = st.prob_qubit(q)
            mask = ((idx >> q) & 1).astype(bool)
            assert abs(p0 - p[~mask].sum()) < 1e-12 and abs(p1 - p[mask].sum()) < 1e-12
        prng = np.random.RandomState(n)
        for label in ["Z" * n, "X" + "I" * (n - 1), "Y" + "Z" * (n - 1)] + \
                ["".join(prng.choice(list("IXYZ"), n)) for _ in range(3)]:
            assert abs(st.expval_pauli(label) - orc.expval_pauli(ref, label)) < 1e-12, label
        u = np.random.RandomState(1).random_sample(20000)
        samples, total = st.sample_all(u)
        assert abs(total - 1.0) < 1e-12
        top = np.argsort(p)[-8:]
        hist = np.bincount(samples, minlength=2 ** n) / 20000.0
        assert np.all(np.abs(hist[top] - p[top]) < 5 * np.sqrt(p[top] / 20000.0) + 1e-3)
        swaps = st.swaps
        full = st.gather_statevector()
        # the sharded sampler is the reference's: searchsorted(cumsum(|psi|^2) / total, u, 'right') over the whole
        # register, running sum chained through the ranks (qasm_simulator.py:213) -- bit for bit
        cdf = (np.abs(full) ** 2).cumsum()
        assert total == cdf[-1]
        assert np.array_equal(samples, (cdf / cdf[-1]).searchsorted(u, side="right"))
        err = float(np.max(np.abs(full - ref)))
        assert err < 1e-12, err
        if rank == 0:
            print("dist ok: world=%d n=%d swaps=%d max-abs err %.2e" % (world, n, swaps, err), flush=True)
    # the BasicAer mirror simulators running SPMD on the sharded CUDA state against the reference's golden outputs
    import golden_io
    import helpers
    from qiskit_sdk_py_b200.dist import distributed_simulator

    names = ["if_statement", "memory_two_registers", "qasm_initial_statevector", "reset_midcircuit",
             "sampler_marginal_subset", "sampler_partial_qubit", "sv_global_phase", "sv_reset_midcircuit", "teleport",
             "transpiled_init_reset_cond_counts", "qv_8_state", "random_12_state", "qft_12_counts", "qv_11_counts",
             "unitary_random", "unitary_global_phase_x4", "random_6_unitary", "unitary_initial_unitary"]
    names += ["midcircuit_8", "sv_midcircuit_10", "qv_16_counts", "random_14_counts"]
    g = int(np.log2(world))
    done = 0
    for name in names:
        case = golden_io.load_case(name)
        nq = case["qobj"]["config"]["n_qubits"] * (2 if case["backend"] == "unitary_simulator" else 1)
        if nq - g < 2:
            continue
        done += 1
        result = helpers.run_case_with(lambda backend: distributed_simulator(backend, chunk_amps=1 << 9), case)
        helpers.compare_result(result, case, amp_tol=1e-10)
    if rank == 0:
        print("spmd ok: %d golden cases through distributed_simulator on %d ranks" % (done, world), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
