#!/usr/bin/env python
"""NVLink peer-memory qubit swap (qsv_swap_bits_p2p) under torchrun on 2+ GPUs: GB/s per GPU and direction as a
function of the swapped local bit, the loads in flight per thread and the grid size.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/p2p_swap_probe.py [n_local]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from qiskit_sdk_py_b200 import _lib  # noqa: E402
from qiskit_sdk_py_b200.dist import ShardedState  # noqa: E402
from qiskit_sdk_py_b200.engine import DeviceState  # noqa: E402

local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
rank, world = dist.get_rank(), dist.get_world_size()
g = world.bit_length() - 1
n_local = int(sys.argv[1]) if len(sys.argv) > 1 else 30
st = ShardedState(n_local + g, DeviceState)
st.init_basis0()
lib = _lib.load()
if len(sys.argv) > 3 and sys.argv[3] == "dense":  # a dense random state instead of |0...0> (are the links data-dependent?)
    from qiskit_sdk_py_b200 import circuits, lower
    ins = [i for i in circuits.quantum_volume_instructions(n_local + g, 3, 5) if i["name"] == "unitary"]
    st.apply_ops([lower.lower_gate(i["name"], i["qubits"], i.get("params")) for i in ins])
    st.restore_identity_layout()
    torch.cuda.synchronize()
    if rank == 0:
        print("dense random state (3 layers of SU(4)), norm2 = %.12f" % st.norm2(), flush=True)
    else:
        st.norm2()


def timed(pairs, ctas, reps=3):
    beta = 0
    for i, (pg, _) in enumerate(pairs):
        beta |= ((rank >> (pg - n_local)) & 1) << i
    lbits = [pl for _, pl in pairs]
    ptrs = [None] * (1 << len(pairs))
    for lam in range(1 << len(pairs)):
        if lam == beta:
            continue
        r = rank
        for i, (pg, _) in enumerate(pairs):
            j = pg - n_local
            r = (r & ~(1 << j)) | (((lam >> i) & 1) << j)
        ptrs[lam] = st._peers[r][0]
    st._stream_barrier()
    st.shard.swap_bits_p2p(lbits, beta, ptrs, ctas)
    st._stream_barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        st._stream_barrier()
        st.shard.swap_bits_p2p(lbits, beta, ptrs, ctas)
    st._stream_barrier()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    k = len(pairs)
    gb = 16.0 * (1 << (n_local - k)) * ((1 << k) - 1) / 1e9
    return float(ms.item()), gb


mode_arg = sys.argv[2] if len(sys.argv) > 2 else ""
quick = mode_arg == "quick"  # chunked vs grid-stride only
if mode_arg == "modes":      # debug library (QSV_LIB_PATH=libqsv_dbg.so): swap vs push-only vs pull-only, all pairs active
    variants = ((0, 4, 0), (1 << 15, 4, 0), (2 << 15, 4, 0))
elif quick:
    variants = ((0, 4, 0), (1 << 14, 4, 1))
else:
    variants = ((0, 4, 0), (1 << 14, 4, 1), (1 << 12, 8, 0), (2 << 12, 2, 0), ((2 << 12) | (1 << 14), 2, 1))
if mode_arg == "bits":       # the swapped local bits as the circuits choose them (k = log2(world) bits per exchange)
    variants = ((0, 4, 0),)
    bit_sets = [[n_local - 1 - j for j in range(g)], [20 - 7 * j for j in range(g)], [9 - 2 * j for j in range(g)],
                [n_local - 1] + [5 + j for j in range(g - 1)]]
    for bits in bit_sets:
        pairs = [(n_local + j, bits[j]) for j in range(g)]
        for ctas in (296,):
            ms, gb = timed(pairs, ctas)
            if rank == 0:
                print("swapped local bits %s ctas %d: %.2f ms  %.1f GB per GPU and direction  %.0f GB/s" % (
                    bits, ctas, ms, gb, gb / ms * 1e3), flush=True)
    variants = ()
for flags, unroll, chunked in variants:
    lib.qsv_debug_set_flags(flags)
    for ctas in ((64, 296) if mode_arg == "modes" else (32, 296) if quick else (16, 32, 64, 148, 296)):
        for bit in ((n_local - 1,) if mode_arg == "modes" else (n_local - 1, n_local // 2, 5)):
            pairs = [(n_local + j, bit - j) for j in range(g)]
            ms, gb = timed(pairs, ctas)
            if rank == 0:
                print("mode %s unroll %d %s ctas %3d swapped local bits %s: %.2f ms  %.1f GB per GPU and direction  %.0f GB/s" % (
                    {0: "swap", 1: "push", 2: "pull"}[(flags >> 15) & 3], unroll, "chunked    " if chunked else "grid-stride", ctas, [pl for _, pl in pairs], ms, gb,
                    gb / ms * 1e3), flush=True)
lib.qsv_debug_set_flags(0)
st.close()
dist.destroy_process_group()
