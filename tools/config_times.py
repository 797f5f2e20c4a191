#!/usr/bin/env python
"""Wall-clock of BASELINE configs 1-3 through the plugin-facing call (qobj in, result out)."""
import copy, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import golden_io
from qiskit_sdk_py_b200 import simulator, circuits

def run(name, backend, qobj):
    sim = simulator.get_simulator(backend)
    sim.run(copy.deepcopy(qobj)); torch.cuda.synchronize()
    t0 = time.perf_counter(); res = sim.run(copy.deepcopy(qobj)); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("%-28s %-22s %8.1f ms  (passes %s, gates %s)" % (name, backend, dt * 1e3, sim.last_stats.get("passes") if hasattr(sim, "last_stats") else "-", sim.last_stats.get("gates") if hasattr(sim, "last_stats") else "-"))
    return res

run("bell (config 1)", "qasm_simulator", circuits.bell())
for name in ("qft_20_product", "qv_20_counts", "qv_24_counts"):
    case = golden_io.load_case(name)
    run(name, case["backend"], case["qobj"])
run("qv_24 statevector", "statevector_simulator", circuits.quantum_volume(24, measure_final=False, shots=1))
run("qv_26 counts (blocked sampler)", "qasm_simulator", circuits.quantum_volume(27, shots=8192))
# QFT at sizes beyond the reference's 24-qubit cap (statevector stays on the device: counts of a product state)
for nq in (28, 30, 32):
    run("qft_%d counts" % nq, "qasm_simulator", circuits.qft(nq, state_seed=3, shots=8192, measure_final=True, seed_simulator=21))
