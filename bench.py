#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 statevector hot path.

Metric (BASELINE.json): gate-pass HBM GB/s & circuit seconds on quantum-volume circuits.
A *step* is one full run of the hot path on one synthetic circuit: initialise |0..0>, apply every
gate of QuantumVolume(n, n, seed=1234) (n*floor(n/2) SU(4) `unitary` ops), draw 8192 shots.

    value  = gates * 32 * 2^n bytes / device time of the step           [GB/s, whole job]
             (32*2^n = one reference gate pass: read + write every complex128 amplitude once,
              SURVEY 8(d); the state is resident in HBM, timed with CUDA events)
    e2e    = the same, measured through the plugin-facing call QasmSimulatorB200.run(qobj) with host
             buffers: Qobj (gate matrices) in, counts out, planning/H2D/D2H inside the timed region
    roofline: the dominant kernel is tile_pass_kernel (one launch = one pass = 32*2^n algorithmic
             bytes); achieved = passes*32*2^n / time of the pass section (CUDA events, live)
    cpu_baseline: the numpy oracle port of the reference (one core: np.einsum is single threaded)
             on a bounded sample of the same circuit family at the reference's 24-qubit cap

N = 1 workload: 33-qubit QV (BASELINE config 4, 128 GiB state) -- the largest single-GPU config.
N > 1: weak scaling, 33 local qubits per GPU (34/35/36 qubits on 2/4/8 GPUs, BASELINE config 5),
one process per GPU, state sharded by the high-order qubits, qubit-swap exchanges over NCCL.

`--impl reference` times the reference's CPU algorithm (oracle port) on a bounded sample.
"""
import argparse
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
REF_GATES_PER_STEP = 6   # bounded sample per reference step (~0.4 s per gate at n=24)
FP64_DMMA_PEAK_TFLOPS = 37.0      # measured: tools/fp64_peak.cu on B200 (DMMA.8x8x4), profiles/r1_fp64_peak_dfma_vs_dmma.txt
TRAFFIC_BYTES_PER_AMP = 31.78     # measured DRAM bytes per amplitude per pass (ncu, n=28; writes still in L2 at kernel end are not counted)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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
def reference_arm(args):
    """The reference's own CPU algorithm (numpy oracle port; np.einsum, one thread) on a bounded
    sample: REF_GATES_PER_STEP SU(4) gates of the same QV family at the reference's 24-qubit cap."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import basicaer_oracle as orc
    from qiskit_sdk_py_b200 import circuits

    n = args.ref_qubits
    ins = circuits.quantum_volume_instructions(n, 2, QV_SEED)
    gates = [i for i in ins if i["name"] == "unitary"][: args.ref_gates]
    state = np.zeros(2 ** n, dtype=complex)
    state[0] = 1.0
    state = state.reshape([2] * n)

    def step(st):
        for g in gates:
            st = orc.apply_unitary(st, g["params"][0], g["qubits"])
        return st

    for _ in range(args.warmup):
        state = step(state)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        state = step(state)
    dt = time.perf_counter() - t0
    bytes_per_gate = 32.0 * 2 ** n
    value = args.steps * len(gates) * bytes_per_gate / dt / 1e9
    sample = "%d SU(4) gates of QuantumVolume(%d, seed=%d) per step (reference cap is 24 qubits)" % (
        len(gates), n, QV_SEED)
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "QuantumVolume gate passes, complex128 statevector, CPU reference algorithm "
                               "(np.einsum per gate, qasm_simulator.py:145-161)", "n_qubits": n,
                   "gates_per_step": len(gates)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(seconds_budget=20.0):
    """Oracle port on this box's host cores (one: np.einsum is single-threaded), bounded sample."""
    from oracle import basicaer_oracle as orc
    from qiskit_sdk_py_b200 import circuits

    n = REF_QUBITS
    ins = [i for i in circuits.quantum_volume_instructions(n, 4, QV_SEED) if i["name"] == "unitary"]
    state = np.zeros(2 ** n, dtype=complex)
    state[0] = 1.0
    state = state.reshape([2] * n)
    state = orc.apply_unitary(state, ins[0]["params"][0], ins[0]["qubits"])  # warm-up / page-in
    done = 0
    t0 = time.perf_counter()
    for g in ins[1:]:
        state = orc.apply_unitary(state, g["params"][0], g["qubits"])
        done += 1
        if time.perf_counter() - t0 > seconds_budget:
            break
    dt = time.perf_counter() - t0
    return {
        "value": done * 32.0 * 2 ** n / dt / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
        "host_cpus": os.cpu_count(), "seconds": dt,
        "sample": "%d SU(4) gate passes of QuantumVolume(%d, seed=%d) (reference cap 24 qubits), "
                  "np.einsum per gate as qasm_simulator.py:145-161" % (done, n, QV_SEED),
    }


# ------------------------------------------------------------------------------------------------
def pick_qubits(requested, torch, reserve=3 << 30):
    free, _total = torch.cuda.mem_get_info()
    n = requested
    while n > 20 and (16 << n) + reserve > free:
        n -= 1
    return n


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
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        from qiskit_sdk_py_b200 import dist as qdist

        return qdist.bench_main(args, METRIC, UNIT, measured_peaks, ClockSampler)

    n = pick_qubits(args.qubits, torch)
    depth = args.depth or n
    tile = args.tile or planner.DEFAULT_TILE_QUBITS
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
    exact = True  # numpy cumsum order at every size (parallel exact scan)
    stream = torch.cuda.current_stream()

    def device_step(ev=None):
        st.init_basis0()
        if ev:
            ev[0].record(stream)
        for p in passes:
            st.apply_pass(p.tile_qubits, p.ops)
        if ev:
            ev[1].record(stream)
        samples, total = st.sample(measured, uniforms, exact_order=exact)
        return samples, total

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
                       "group_variant_passes": sum(i["threads"] == 768 for i in infos)},
        "peak_source": "measured on B200 with tools/fp64_peak.cu (profiles/r1_fp64_peak_dfma_vs_dmma.txt)",
        "flops": "executed: 32 flop/amplitude per lone 2-qubit op (2 DMMA.8x8x4 per 32 amplitudes), 49 per fused "
                 "3-qubit block in the 3-multiplication form (6 DMMA + 2 DADD per 64 amplitudes); nominal: 32 per "
                 "SU(4) gate",
    }

    # ---- end-to-end leg through the plugin-facing call, host buffers in and out ------------------
    e2e = None
    if not args.no_e2e:
        del st
        torch.cuda.empty_cache()
        sim = simulator.QasmSimulatorB200(tile_qubits=tile, max_qubits=max(n, 33))
        import copy

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

    # ---- second workload family of BASELINE's metric (informational, outside the timed headline): QFT on the
    # same register, in the simulator basis, through the same plugin-facing call ---------------------
    qft = None
    if not args.no_e2e:
        qft_qobj = circuits.qft(n, state_seed=3, shots=SHOTS, measure_final=True, seed_simulator=SIM_SEED)
        n_ins = sum(1 for i in qft_qobj["experiments"][0]["instructions"] if i["name"] not in ("measure", "barrier"))
        sim.run(copy.deepcopy(qft_qobj))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sim.run(copy.deepcopy(qft_qobj))
        torch.cuda.synchronize()
        qft_s = time.perf_counter() - t0
        qft = {"workload": "QFT(%d) with final swaps on a product state, basis gates u1/u2/u3/cx, %d-shot sampling" % (n, SHOTS),
               "instructions": n_ins, "passes": sim.last_stats.get("passes"), "circuit_s": qft_s,
               "reference_equivalent_GBps": n_ins * bytes_per_gate_pass / qft_s / 1e9,
               "note": "the reference makes one einsum pass per instruction; peephole merging, diagonal tables and "
                       "diagonal tails bring the circuit down to `passes` passes over HBM"}

    cpu = None
    if not args.no_cpu_baseline:
        cpu = cpu_baseline_leg()

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": "QuantumVolume(%d, depth %d, seed %d): %d SU(4) gates + %d-shot sampling, complex128 "
                        "statevector (%.1f GiB) on one B200 -- BASELINE config 4" % (
                            n, depth, QV_SEED, n_gates, SHOTS, 16 * 2 ** n / 2 ** 30),
            "n_qubits": n, "depth": depth, "gates": n_gates, "shots": SHOTS, "tile_qubits": tile,
            "passes": len(passes), "gates_per_pass": n_gates / len(passes),
            "circuit_s": step_ms * 1e-3,
            "l2": "state (%.1f GiB) is far larger than the 126 MB L2; no flush needed" % (16 * 2 ** n / 2 ** 30),
            "sampler": "exact-order" if exact else "blocked-parallel",
        },
        "roofline": {
            "bound": "hbm", "kernel": "tile_pass_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": TRAFFIC_BYTES_PER_AMP * 2 ** n, "peak_source": peak_src,
            "traffic_source": "ncu --set full at n=28: dram__bytes_read 4.295 GB + write 4.236 GB per launch = "
                              "31.8 B/amplitude (profiles/r1_tile_pass_kernel_ncu_full_summary.txt), scaled to 2^n",
            "algorithmic_bytes_per_launch": bytes_per_gate_pass, "launch_ms": per_launch_ms,
            "launches_per_step": len(passes), "pass_section_ms": pass_ms,
            # the fused pass is FP64-bound, not HBM-bound: the same launches against the measured DMMA peak
            "fp64": fp64_info,
        },
        "cpu_baseline": cpu,
        "e2e": e2e,
        "qft": qft,
        "gpu_launches": int(launches),
        "clocks": clock_info,
        "check": {"sample_total_prob": total, "max_bin_mod4096": counts_top},
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
