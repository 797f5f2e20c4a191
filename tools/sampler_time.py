import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from qiskit_sdk_py_b200 import circuits, lower
from qiskit_sdk_py_b200.engine import DeviceState
for n in (24, 28, 30):
    st = DeviceState(n); st.init_basis0()
    ops = [lower.lower_gate(i["name"], i["qubits"], i.get("params")) for i in circuits.quantum_volume_instructions(n, 3, 7)]
    st.apply_ops(ops)
    u = np.random.RandomState(1).random_sample(8192)
    for exact in (True, False):
        st.sample(list(range(n)), u, exact_order=exact); torch.cuda.synchronize()
        t0 = time.perf_counter(); a, tot = st.sample(list(range(n)), u, exact_order=exact); torch.cuda.synchronize()
        print(n, "exact" if exact else "blocked", "%.2f ms" % ((time.perf_counter() - t0) * 1e3), tot)
    e, _ = st.sample(list(range(n)), u, exact_order=True); b, _ = st.sample(list(range(n)), u, exact_order=False)
    print("  mismatching shots exact vs blocked:", int((e != b).sum()))
    del st; torch.cuda.empty_cache()
