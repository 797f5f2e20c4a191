#!/usr/bin/env python
"""A random basis-gate circuit (u1/u2/u3/rz/sx/x/cx layers, circuits.random_circuit) at n qubits: op counts after
merge_ops, passes, time.  python tools/random_probe.py [n_qubits] [depth] [tile_qubits]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qiskit_sdk_py_b200 import circuits, lower, planner
from qiskit_sdk_py_b200.engine import DeviceState

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 20
tile = int(sys.argv[3]) if len(sys.argv) > 3 else planner.DEFAULT_TILE_QUBITS
ins = circuits.random_basis_instructions(n, depth, 7)
ops = [o for o in (lower.lower_gate(i["name"], i["qubits"], i.get("params")) for i in ins) if o is not None]
merged = planner.merge_ops(ops)
passes = planner.plan_passes(merged, n, tile, tails=True)
kinds = {}
for o in merged:
    kinds[o.kind] = kinds.get(o.kind, 0) + 1
kernel_ops = sum(DeviceState.pass_info(n, p.tile_qubits, p.ops)["ops"] for p in passes)
st = DeviceState(n, tile_qubits=tile)
st.init_basis0()
for p in passes:
    st.apply_pass(p.tile_qubits, p.ops, tail=p.tail)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for p in passes:
    st.apply_pass(p.tile_qubits, p.ops, tail=p.tail)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("random circuit n=%d depth=%d tile=%d: %d instructions -> %d merged ops %s -> %d kernel ops in %d passes (%d tail gates): "
      "%.1f ms, %.2f ms/pass, reference-equivalent %.0f GB/s" % (
          n, depth, tile, len(ops), len(merged), kinds, kernel_ops, len(passes), sum(len(p.tail) for p in passes), ms,
          ms / len(passes), len(ops) * 32 * 2 ** n / ms / 1e6))
