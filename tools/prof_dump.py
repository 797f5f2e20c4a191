#!/usr/bin/env python
"""Per-phase clock breakdown of the warp-group pass kernel (needs the profiling build: QSV_LIB_PATH=.../libqsv_dbg.so).
Runs the first passes of QuantumVolume(n) and prints, per pass, the mean clocks per tile spent by a compute warp
(wait for the tile / gates / loop overhead) and by the group's TMA producer warp (wait for the gates / issue stores /
wait for the stores to read the buffer / issue loads).  Usage: prof_dump.py [n_qubits] [n_passes]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from qiskit_sdk_py_b200 import _lib, circuits, lower, planner  # noqa: E402
from qiskit_sdk_py_b200.engine import DeviceState  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
n_passes = int(sys.argv[2]) if len(sys.argv) > 2 else 12
ins = [i for i in circuits.quantum_volume_instructions(n, n, 1234) if i["name"] == "unitary"]
ops = [lower.lower_gate(i["name"], i["qubits"], i.get("params")) for i in ins]
passes = planner.plan_passes(ops, n, planner.DEFAULT_TILE_QUBITS)
st = DeviceState(n)
st.init_basis0()
lib = _lib.load()
SLOTS = 20
ROLES = 9
words = 148 * 3 * ROLES * SLOTS
buf = (ctypes.c_uint64 * words)()
tiles_per_group = (1 << (n - 11)) / (148 * 3)
print("n=%d: %d passes, %.0f tiles per group; clocks per tile" % (n, len(passes), tiles_per_group))
print("%4s %5s %4s | compute warp0: wait_tile gates other | warp7: wait_tile gates other | producer: wait_gates st_issue "
      "wait_read ld_issue | ms" % ("pass", "gates", "ops"))
for k, p in enumerate(passes[:n_passes]):
    info = DeviceState.pass_info(n, p.tile_qubits, p.ops)
    for _ in range(2):
        st.apply_pass(p.tile_qubits, p.ops)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    st.apply_pass(p.tile_qubits, p.ops)
    e1.record()
    torch.cuda.synchronize()
    _lib.check(lib.qsv_debug_profile(buf, words), "qsv_debug_profile")
    a = np.frombuffer(buf, dtype=np.uint64).reshape(148, 3, ROLES, SLOTS).astype(np.float64) / tiles_per_group
    m = a.mean(axis=(0, 1))
    print("%4d %5d %4d | %7.0f %7.0f %6.0f | %7.0f %7.0f %6.0f | %7.0f %7.0f %7.0f %7.0f | %.3f" % (
        k, len(p.ops), info["ops"], m[0, 0], m[0, 1], m[0, 2], m[7, 0], m[7, 1], m[7, 2], m[8, 0], m[8, 1], m[8, 2],
        m[8, 3], e0.elapsed_time(e1)))
    print("       gates per warp 0..7: %s | wait_tile: %s" % (" ".join("%6.0f" % v for v in m[:8, 1]),
                                                                " ".join("%5.0f" % v for v in m[:8, 0])))
    kinds = DeviceState.pass_ops(n, p.tile_qubits, p.ops) if hasattr(DeviceState, "pass_ops") else None
    print("       warp0 per op: barrier %5.0f | %s" % (m[0, 4], " ".join("%5.0f" % v for v in m[0, 5:5 + info["ops"]])))
    print("       warp7 per op: barrier %5.0f | %s" % (m[7, 4], " ".join("%5.0f" % v for v in m[7, 5:5 + info["ops"]])))
