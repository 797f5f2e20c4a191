#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 statevector hot path.

Metric (BASELINE.json): gate-pass HBM GB/s & circuit seconds on quantum-volume / QFT / random circuits.
A *step* is one full run of the hot path on one synthetic circuit: initialise |0..0>, apply every
gate of QuantumVolume(n, n, seed=1234) (n*floor(n/2) SU(4) `unitary` ops), draw 8192 shots.

    value  = gates * 32 * 2^n bytes / device time of the step           [GB/s, whole job]
             (32*2^n = one reference gate pass: read + write every complex128 amplitude once,
              SURVEY 8(d); the state is resident in HBM, timed with CUDA events)
    e2e    = the same, measured through the plugin-facing call QasmSimulatorB200.run(qobj) with host
             buffers: Qobj (gate matrices) in, counts out, planning/H2D/D2H inside the timed region
    roofline: the dominant kernel is tile_pass_kernel (one launch = one pass = 32*2^n algorithmic
             bytes); achieved = passes*32*2^n / time of the pass section (CUDA events, live)
    cpu_baseline: the reference's own QasmSimulatorPy.run (np.einsum per gate, one core) on a bounded
             sample of the same circuit family at the reference's 24-qubit cap
    qft / random: the other two circuit families of BASELINE's metric, with per-pass rooflines

N = 1 workload: 33-qubit QV (BASELINE config 4, 128 GiB state) -- the largest single-GPU config.
N > 1: weak scaling, 33 local qubits per GPU (34/35/36 qubits on 2/4/8 GPUs, BASELINE config 5),
one process per GPU, state sharded by the high-order qubits, qubit-swap exchanges over NCCL; before the
timed region the golden fixtures run through the sharded simulators as a parity gate (`check.parity`).

`--impl reference` times the reference's own CPU implementation (QasmSimulatorPy.run from the copy under
baseline/_ref; the numpy oracle port if that cannot be imported) on a bounded sample per step.
"""
import argparse
import copy
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "qv_gate_pass_effective_bandwidth"
UNIT = "GB/s"
SHOTS = 8192
QV_SEED = 1234
SIM_SEED = 42
REF_QUBITS = 24          # reference cap (qasm_simulator.py:64)
REF_GATES_PER_STEP = 24  # bounded CPU sample per step: every 12th SU(4) of QuantumVolume(24) (~0.3-0.9 s per gate)
RANDOM_DEPTH = 20
FP64_DMMA_PEAK_TFLOPS = 37.0      # measured: tools/fp64_peak.cu on B200 (DMMA.8x8x4), profiles/r1_fp64_peak_dfma_vs_dmma.txt
NCU_TRAFFIC_FILE = os.path.join("profiles", "r2_tile_pass_ncu_dram.json")  # written by tools/ncu_summary.py from an ncu CSV


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(n):
    """DRAM bytes per launch of the pass kernel from the committed ncu capture (dram__bytes_read.sum +
    dram__bytes_write.sum of one `ncu --set full` launch), scaled by 2^(n - n_capture) when the sizes differ."""
    path = os.path.join(ROOT, NCU_TRAFFIC_FILE)
    if not os.path.exists(path):
        return None, "no ncu capture committed (%s missing)" % NCU_TRAFFIC_FILE
    with open(path) as fh:
        rec = json.load(fh)
    per_launch = float(rec["dram_bytes_read"]) + float(rec["dram_bytes_write"])
    scale = 2.0 ** (n - int(rec["n_qubits"]))
    return per_launch * scale, "%s: dram__bytes_read %.4g + dram__bytes_write %.4g per launch at n=%d (%s)%s" % (
        NCU_TRAFFIC_FILE, rec["dram_bytes_read"], rec["dram_bytes_write"], rec["n_qubits"], rec.get("source", "ncu --set full"),
        "" if scale == 1.0 else ", scaled by 2^%d" % (n - int(rec["n_qubits"])))


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.QUERY,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": max(smax) if smax else None,
            "power_w_max": max(power) if power else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


# ------------------------------------------------------------------------------------------------
def workload_config(n, depth, tile, world=1, n_local=None):
    """Description of the workload -- the same dict in the B200 arm and in the reference arm (host-only: the
    planner runs without a GPU)."""
    from qiskit_sdk_py_b200 import circuits, lower, planner

    ins = [i for i in circuits.quantum_volume_instructions(n, depth, QV_SEED) if i["name"] == "unitary"]
    cfg = {
        "workload": "QuantumVolume(%d, depth %d, seed %d): %d SU(4) gates + %d-shot sampling, complex128 statevector "
                    "(%.1f GiB)%s -- BASELINE config %d" % (
                        n, depth, QV_SEED, len(ins), SHOTS, 16 * 2 ** n / 2 ** 30,
                        " on one B200" if world == 1 else " sharded over %d B200 (%d local qubits, %.1f GiB per GPU)" % (
                            world, n_local, 16 * 2 ** n_local / 2 ** 30), 4 if world == 1 else 5),
        "n_qubits": n, "depth": depth, "gates": len(ins), "shots": SHOTS, "tile_qubits": tile,
        "l2": "state (%.1f GiB per GPU) is far larger than the 126 MB L2; no flush needed" % (
            16 * 2 ** (n_local or n) / 2 ** 30),
        "sampler": "exact-order (numpy cumsum order, bit-exact seeded counts)",
    }
    if world == 1:
        ops = [lower.lower_gate(i["name"], i["qubits"], i.get("params")) for i in ins]
        passes = planner.plan_passes(ops, n, tile)
        cfg["passes"] = len(passes)
        cfg["gates_per_pass"] = len(ins) / len(passes)
    else:
        cfg["local_qubits"] = n_local
        cfg["parallelism"] = ("state sharded by the %d high-order qubits; qubit swaps = one in-place all-to-all kernel per "
                              "rank over NVLink peer memory (CUDA IPC), NCCL only for barriers and scalar reductions"
                              % int(round(np.log2(world))))
    return cfg


# ---- the reference's CPU implementation on a bounded sample ----------------------------------------
def cpu_sample_qobj(n=REF_QUBITS, n_gates=REF_GATES_PER_STEP):
    """`n_gates` SU(4) gates spread evenly over QuantumVolume(n, n, seed) (every k-th gate of the circuit, so the
    sample sees the same mix of qubit pairs -- einsum's speed depends on which axes a gate touches), no measurement."""
    from qiskit_sdk_py_b200 import circuits

    ins = [i for i in circuits.quantum_volume_instructions(n, n, QV_SEED) if i["name"] == "unitary"]
    stride = max(1, len(ins) // n_gates)
    pick = ins[::stride][:n_gates]
    exp = circuits.make_experiment("qv_sample_%d" % n, n, pick, memory_slots=0)
    return circuits.make_qobj([exp], shots=1, seed_simulator=SIM_SEED), len(pick), stride


def reference_runner():
    """(kind, run(qobj_dict) -> result) -- the reference's own QasmSimulatorPy when its package can be imported
    (baseline/_ref on the GPU box, /root/reference in the build container), else the numpy oracle port."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_env"))
        import warnings

        import ref_import

        ref_import.import_reference()
        warnings.simplefilter("ignore")
        from qiskit import BasicAer
        from qiskit.qobj import QasmQobj

        backend = BasicAer.get_backend("qasm_simulator")

        def run(qobj_dict):
            return backend.run(QasmQobj.from_dict(copy.deepcopy(qobj_dict))).result()

        return "reference", run, "QasmSimulatorPy.run(qobj) of the unmodified reference (%s)" % ref_import.REFERENCE_ROOT
    except Exception as exc:  # the reference package is not importable here: the port restates the same numpy calls
        from oracle import basicaer_oracle as orc

        sim = orc.OracleQasmSimulator()
        return "port", (lambda qobj_dict: sim.run(copy.deepcopy(qobj_dict))), "oracle port (reference import failed: %s)" % exc


def cpu_reference_leg(steps, warmup, n=REF_QUBITS, n_gates=REF_GATES_PER_STEP):
    kind, run, how = reference_runner()
    qobj, picked, stride = cpu_sample_qobj(n, n_gates)
    warm_qobj, _, _ = cpu_sample_qobj(n, 2)
    for _ in range(max(1, warmup) if warmup else 0):
        run(warm_qobj)  # page-in / allocator warm-up on a 2-gate sample (a full sample step costs ~10 s)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        run(qobj)
        times.append(time.perf_counter() - t0)
    dt = sum(times)
    value = steps * picked * 32.0 * 2 ** n / dt / 1e9
    return {
        "value": value, "unit": UNIT, "cores": 1, "kind": kind, "host_cpus": os.cpu_count(), "seconds": dt,
        "ms_per_step": dt / steps * 1e3,
        "sample": "%d SU(4) gates per step = every %dth gate of QuantumVolume(%d, depth %d, seed %d) (the reference caps "
                  "at 24 qubits), shots=1, through %s; np.einsum is single-threaded" % (picked, stride, n, n, QV_SEED, how),
    }


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from qiskit_sdk_py_b200 import planner

    world = max(1, args.gpus)
    n = args.qubits + int(round(np.log2(world)))
    cfg = workload_config(n, args.depth or n, args.tile or planner.DEFAULT_TILE_QUBITS, world, args.qubits)
    leg = cpu_reference_leg(args.steps, args.warmup, args.ref_qubits, args.ref_gates)
    line = {
        "impl": "reference",
        "metric": METRIC, "value": leg["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": leg["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample", "host_cpus", "seconds")},
        "e2e": {"value": leg["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def pick_qubits(requested, torch, reserve=3 << 30):
    free, _total = torch.cuda.mem_get_info()
    n = requested
    while n > 20 and (16 << n) + reserve > free:
        n -= 1
    return n


def timed_steps(st, steps, torch, n):
    """Run the steps of DeviceState.plan_flush (pass-kernel launches and qubit-permutation launches) one by one with CUDA
    events; returns per-step ms."""
    stream = torch.cuda.current_stream()
    events = []
    for step in steps:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        st.run_step(step)
        e1.record(stream)
        events.append((e0, e1))
    torch.cuda.synchronize()
    return [a.elapsed_time(b) for a, b in events]


def family_block(name, instructions, n, tile, st, torch, peak):
    """Per-pass roofline of another circuit family (QFT, random basis-gate circuit) on the resident state: the launches
    DeviceState.apply_ops makes for it (fused passes; one in-place permutation launch for the swaps taken out of the
    stream), each a read + write of the whole state."""
    from qiskit_sdk_py_b200 import lower, planner

    ops = [o for o in (lower.lower_gate(i["name"], i["qubits"], i.get("params")) for i in instructions) if o is not None]
    steps = st.plan_flush(ops)
    passes = [what for kind, what in steps if kind == "pass"]
    st.init_basis0()
    timed_steps(st, steps, torch, n)  # warm-up
    ms = timed_steps(st, steps, torch, n)
    per_pass_gbs = [32.0 * 2 ** n / (t * 1e-3) / 1e9 for t in ms]
    return {
        "workload": name, "instructions": len(ops), "merged_ops": len(planner.merge_ops(ops)), "passes": len(steps),
        "swap_kernel_launches": len(steps) - len(passes),
        "tail_gates": sum(len(p.tail) for p in passes), "pass_section_ms": sum(ms),
        "reference_equivalent_GBps": len(ops) * 32.0 * 2 ** n / (sum(ms) * 1e-3) / 1e9,
        "roofline": {"bound": "hbm", "kernel": "tile_pass_kernel", "unit": "GB/s", "peak": peak,
                     "achieved": len(steps) * 32.0 * 2 ** n / (sum(ms) * 1e-3) / 1e9,
                     "frac": len(steps) * 32.0 * 2 ** n / (sum(ms) * 1e-3) / 1e9 / peak,
                     "per_pass_ms": [round(t, 3) for t in ms],
                     "per_pass_kind": ["pass" if kind == "pass" else "swap_pairs" for kind, _ in steps],
                     "per_pass_frac": [round(g / peak, 3) for g in per_pass_gbs],
                     "worst_pass_frac": min(per_pass_gbs) / peak},
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--qubits", type=int, default=33, help="qubits per GPU (local qubits)")
    ap.add_argument("--depth", type=int, default=0, help="QV depth (0 = number of qubits)")
    ap.add_argument("--tile", type=int, default=0, help="tile qubits (0 = engine default)")
    ap.add_argument("--ref-qubits", type=int, default=REF_QUBITS)
    ap.add_argument("--ref-gates", type=int, default=REF_GATES_PER_STEP)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-families", action="store_true", help="skip the QFT / random-circuit blocks")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the golden parity gate")
    args = ap.parse_args()

    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on
    # fd 1), so fd 1 is pointed at stderr for the whole run and the JSON line goes to the saved real stdout.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout

    if args.impl == "reference":
        reference_arm(args)
        return

    import torch

    from qiskit_sdk_py_b200 import _lib, circuits, lower, planner, simulator
    from qiskit_sdk_py_b200.engine import DeviceState

    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        from qiskit_sdk_py_b200 import dist as qdist

        return qdist.bench_main(args, sys.modules[__name__])

    n = pick_qubits(args.qubits, torch)
    depth = args.depth or n
    tile = args.tile or planner.DEFAULT_TILE_QUBITS
    cfg = workload_config(n, depth, tile)
    qobj = circuits.quantum_volume(n, depth, QV_SEED, shots=SHOTS, seed_simulator=SIM_SEED)
    instructions = qobj["experiments"][0]["instructions"]
    gate_ins = [i for i in instructions if i["name"] == "unitary"]
    n_gates = len(gate_ins)
    bytes_per_gate_pass = 32.0 * 2 ** n

    # ---- device-resident leg: `value` and the roofline of the dominant kernel --------------------
    ops = [lower.lower_gate(i["name"], i["qubits"], i.get("params")) for i in gate_ins]
    passes = planner.plan_passes(ops, n, tile)
    st = DeviceState(n, tile_qubits=tile)
    uniforms = np.random.RandomState(SIM_SEED).random_sample(SHOTS)
    measured = list(range(n))
    stream = torch.cuda.current_stream()

    def device_step(ev=None):
        st.init_basis0()
        if ev:
            ev[0].record(stream)
        for p in passes:
            st.apply_pass(p.tile_qubits, p.ops)
        if ev:
            ev[1].record(stream)
        return st.sample(measured, uniforms, exact_order=True)

    for _ in range(args.warmup):
        device_step()
    torch.cuda.synchronize()
    clocks = ClockSampler(local_rank)
    clocks.start()
    launches0 = _lib.launch_count()
    pass_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                   for _ in range(args.steps)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t_begin.record(stream)
    for k in range(args.steps):
        samples, total = device_step(pass_events[k])
    t_end.record(stream)
    torch.cuda.synchronize()
    launches = _lib.launch_count() - launches0
    clock_info = clocks.stop()
    step_ms = t_begin.elapsed_time(t_end) / args.steps
    pass_ms = sum(a.elapsed_time(b) for a, b in pass_events) / args.steps
    value = n_gates * bytes_per_gate_pass / (step_ms * 1e-3) / 1e9
    peak, peak_src = measured_peaks()
    per_launch_ms = pass_ms / len(passes)
    achieved = bytes_per_gate_pass / (per_launch_ms * 1e-3) / 1e9
    assert abs(total - 1.0) < 1e-9, total
    counts_top = int(np.bincount(samples % 4096).max())
    traffic, traffic_src = measured_traffic(n)

    # FP64 work of the pass launches.  `nominal`: 32 flop per amplitude per SU(4) gate (a 4x4 complex matrix on
    # groups of 4 amplitudes).  `executed`: what the kernel issues after gate fusion -- a lone 2-qubit op is 2
    # DMMA.8x8x4 per 32 amplitudes (32 flop/amplitude); a fused 3-qubit block (two or more gates) in the
    # 3-multiplication form is 6 DMMA + 2 DADD per 64 amplitudes (49 flop/amplitude).
    infos = [DeviceState.pass_info(n, p.tile_qubits, p.ops) for p in passes]
    n_blocks = sum(i["fused3"] for i in infos)
    n_single = sum(i["ops"] - i["fused3"] for i in infos)
    flops_nominal = n_gates * 32.0 * 2 ** n
    flops_executed = (n_single * 32.0 + n_blocks * 49.0) * 2 ** n
    fp64_info = {
        "achieved_tflops": flops_executed / (pass_ms * 1e-3) / 1e12, "peak_tflops": FP64_DMMA_PEAK_TFLOPS,
        "frac": flops_executed / (pass_ms * 1e-3) / 1e12 / FP64_DMMA_PEAK_TFLOPS,
        "nominal_tflops": flops_nominal / (pass_ms * 1e-3) / 1e12,
        # the pass runs at the 1 kW power cap: DMMA peak at the SM clock measured during the timed region
        # (148 SMs x 4 sub-partitions x one DMMA.8x8x4 = 512 flop per 16 cycles)
        "peak_tflops_at_measured_clock": (148 * 128 * clock_info["sm_mhz"] * 1e6 / 1e12
                                          if clock_info.get("sm_mhz") else None),
        "frac_at_measured_clock": (flops_executed / (pass_ms * 1e-3) / (148 * 128 * clock_info["sm_mhz"] * 1e6)
                                   if clock_info.get("sm_mhz") else None),
        "kernel_ops": {"fused_3qubit_blocks": n_blocks, "2qubit_ops": n_single,
                       "group_variant_passes": sum(i["threads"] == 864 for i in infos)},
        "peak_source": "measured on B200 with tools/fp64_peak.cu (profiles/r1_fp64_peak_dfma_vs_dmma.txt)",
        "flops": "executed: 32 flop/amplitude per lone 2-qubit op (2 DMMA.8x8x4 per 32 amplitudes), 49 per fused "
                 "3-qubit block in the 3-multiplication form (6 DMMA + 2 DADD per 64 amplitudes); nominal: 32 per "
                 "SU(4) gate",
    }

    # ---- the other circuit families of BASELINE's metric, per-pass rooflines on the same resident state ----
    qft = rnd = None
    if not args.no_families:
        qft_ins = [i for i in circuits.qft(n, state_seed=3)["experiments"][0]["instructions"]
                   if i["name"] not in ("measure", "barrier")]
        qft = family_block("QFT(%d) with final swaps on a product state, basis gates u1/u2/u3/cx" % n, qft_ins, n, tile, st,
                           torch, peak)
        rnd_ins = circuits.random_basis_instructions(n, RANDOM_DEPTH, 7)
        rnd = family_block("random basis-gate circuit (u1/u2/u3/rz/sx/x/cx + 2-qubit unitaries), %d qubits, depth %d, seed 7"
                           % (n, RANDOM_DEPTH), rnd_ins, n, tile, st, torch, peak)

    # ---- end-to-end leg through the plugin-facing call, host buffers in and out ------------------
    e2e = None
    if not args.no_e2e:
        del st
        torch.cuda.empty_cache()
        sim = simulator.QasmSimulatorB200(tile_qubits=tile, max_qubits=max(n, 33))
        sim.run(copy.deepcopy(qobj))  # warm-up (allocation, attribute set-up)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = sim.run(copy.deepcopy(qobj))
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / args.steps
        counts = res["results"][0]["data"]["counts"]
        assert sum(counts.values()) == SHOTS
        h2d = sum(np.asarray(i["params"][0]).nbytes for i in gate_ins) + SHOTS * 8
        d2h = SHOTS * 8 + 8
        e2e = {"value": n_gates * bytes_per_gate_pass / e2e_s / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "circuit_s": e2e_s,
               "api": "QasmSimulatorB200.run(qobj dict) -> counts"}
        if qft is not None:  # the QFT family end to end through the same call
            qft_qobj = circuits.qft(n, state_seed=3, shots=SHOTS, measure_final=True, seed_simulator=SIM_SEED)
            sim.run(copy.deepcopy(qft_qobj))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            sim.run(copy.deepcopy(qft_qobj))
            torch.cuda.synchronize()
            qft["e2e_circuit_s"] = time.perf_counter() - t0
            rnd_qobj = circuits.random_circuit(n, RANDOM_DEPTH, 7, shots=SHOTS, seed_simulator=SIM_SEED)
            sim.run(copy.deepcopy(rnd_qobj))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            sim.run(copy.deepcopy(rnd_qobj))
            torch.cuda.synchronize()
            rnd["e2e_circuit_s"] = time.perf_counter() - t0

    cpu = None
    if not args.no_cpu_baseline:
        cpu = cpu_reference_leg(1, 1, args.ref_qubits, args.ref_gates)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "circuit_s": step_ms * 1e-3,
        "roofline": {
            "bound": "hbm", "kernel": "tile_pass_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src, "traffic_source": traffic_src,
            "algorithmic_bytes_per_launch": bytes_per_gate_pass, "launch_ms": per_launch_ms,
            "launches_per_step": len(passes), "pass_section_ms": pass_ms,
            # the fused pass is FP64-bound, not HBM-bound: the same launches against the measured DMMA peak
            "fp64": fp64_info,
        },
        "cpu_baseline": cpu,
        "e2e": e2e,
        "qft": qft,
        "random_circuit": rnd,
        "gpu_launches": int(launches),
        "clocks": clock_info,
        "check": {"sample_total_prob": total, "max_bin_mod4096": counts_top},
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
