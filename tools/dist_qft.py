#!/usr/bin/env python
"""QFT on a sharded state (one process per GPU, torchrun): python -m torch.distributed.run --nproc-per-node N
tools/dist_qft.py [local_qubits].  Prints time, local passes and qubit exchanges, and checks the DFT column."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from qiskit_sdk_py_b200 import circuits, lower
from qiskit_sdk_py_b200.dist import ShardedState, DEFAULT_CHUNK_AMPS
from qiskit_sdk_py_b200.engine import DeviceState

world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
g = int(round(math.log2(world)))
n_local = int(sys.argv[1]) if len(sys.argv) > 1 else 33
n = n_local + g
x = ((1 << (n - 1)) | 0x2D5 | (1 << (n // 2))) & ((1 << n) - 1)
ins = [{"name": "x", "qubits": [q]} for q in range(n) if (x >> q) & 1] + circuits.qft(n)["experiments"][0]["instructions"]
ops = [o for o in (lower.lower_gate(i["name"], i["qubits"], i.get("params")) for i in ins) if o is not None]
st = ShardedState(n, DeviceState, chunk_amps=min(DEFAULT_CHUNK_AMPS, 1 << (n_local - 1)))

def step():
    st.init_basis0()
    st.apply_ops(ops)
    st.restore_identity_layout()

step()
torch.cuda.synchronize(); dist.barrier()
p0, s0, f0 = st.shard.passes_launched, st.swaps, st.fused_exchanges
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); torch.cuda.synchronize(); dist.barrier()
ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
# DFT column check on this rank's slab: global index y = rank * 2^n_local + local index
ys_local = [0, 1, (1 << n_local) - 1] + [int(v) for v in np.random.RandomState(rank).randint(0, 2 ** 31, 253) % (1 << n_local)]
idx = torch.tensor([2 * y for y in ys_local] + [2 * y + 1 for y in ys_local], dtype=torch.int64, device=st.shard.state.device)
vals = st.shard.state.index_select(0, idx).cpu().numpy()
got = vals[: len(ys_local)] + 1j * vals[len(ys_local):]
ys = [(rank << n_local) + y for y in ys_local]
want = np.exp(2j * np.pi * np.array([((x * y) % (1 << n)) / float(1 << n) for y in ys])) / np.sqrt(float(1 << n))
err = torch.tensor([float(np.max(np.abs(got - want)) * np.sqrt(float(1 << n)))], device="cuda", dtype=torch.float64)
dist.all_reduce(err, op=dist.ReduceOp.MAX)
if rank == 0:
    print("QFT-%d on %d GPUs (%d local qubits): %d instructions, %.1f ms, %d local passes, %d qubit swaps in %d exchanges, "
          "max relative amplitude error vs DFT column %.2e" % (n, world, n_local, len(ops), float(ms.item()),
          st.shard.passes_launched - p0, st.swaps - s0, st.fused_exchanges - f0, float(err.item())))
dist.destroy_process_group()
