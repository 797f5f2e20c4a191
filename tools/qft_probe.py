#!/usr/bin/env python
"""Per-pass times of a QFT (merge_ops + diagonal tails): python tools/qft_probe.py [n_qubits]."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qiskit_sdk_py_b200 import circuits, lower, planner
from qiskit_sdk_py_b200.engine import DeviceState

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
ins = circuits.qft(n)["experiments"][0]["instructions"]
ops = [o for o in (lower.lower_gate(i["name"], i["qubits"], i.get("params")) for i in ins) if o is not None]
merged = planner.merge_ops(ops)
st = DeviceState(n)
steps = st.plan_flush(ops)  # what DeviceState.apply_ops launches: fused passes + the permutation launch for the swaps
passes = steps
st.init_basis0()
for step in steps:  # warm-up
    st.run_step(step)
torch.cuda.synchronize()
total = 0.0
for step in steps:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    st.run_step(step)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    total += ms
    if step[0] == "swap":
        print("swap_qubit_pairs: %2d pairs %s : %7.3f ms" % (len(step[1]), step[1], ms))
        continue
    p = step[1]
    info = DeviceState.pass_info(n, p.tile_qubits, p.ops)
    kinds = {}
    for o in p.ops:
        kinds[o.kind] = kinds.get(o.kind, 0) + 1
    print("pass: %2d gates %s -> %2d kernel ops, %d stages, tail %3d : %7.3f ms" % (len(p.ops), kinds, info["ops"], info["stages"], len(p.tail), ms))
print("QFT-%d: %d instructions -> %d merged ops -> %d passes, %.1f ms (%.1f GB/s per pass of 32*2^n B)" % (
    n, len(ops), len(merged), len(passes), total, len(passes) * 32 * 2 ** n / total / 1e6))
