"""The C-ABI library loads and exports every symbol include/qsv.h declares (no compute calls: there is
no GPU in the CPU test environment)."""
import ctypes
import os
import re

import pytest

from qiskit_sdk_py_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "qsv.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qsv_[a-z0-9_]+)\s*\(", text)))


def test_library_built():
    assert os.path.exists(_lib.LIB_PATH), "run `python __graft_entry__.py` to build libqsv.so"


def test_every_declared_symbol_is_exported_and_bound():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 20
    for name in syms:
        assert hasattr(lib, name), "libqsv.so does not export %s" % name
        assert name in _lib.SIGNATURES, "no ctypes signature for %s" % name
    assert set(_lib.SIGNATURES) == set(syms)


def test_version_and_error_string():
    lib = _lib.load()
    assert lib.qsv_version() == 100
    assert isinstance(lib.qsv_last_error(), bytes)
    assert lib.qsv_reduce_workspace_bytes() > 0
    assert lib.qsv_sample_workspace_bytes(20, 1024) > 8 * 1024
    assert lib.qsv_marginal_workspace_bytes(20, 3) >= 8 * 8


def test_argument_validation_without_gpu():
    """Argument errors are reported before any CUDA call."""
    lib = _lib.load()
    rc = lib.qsv_apply_pass(None, 4, None, 0, None, 0, None)
    assert rc == _lib.QSV_EINVAL
    assert b"qsv_apply_pass" in lib.qsv_last_error()
    assert lib.qsv_prob_qubit(None, 4, 0, None, None, None) == _lib.QSV_EINVAL
    import ctypes

    pairs = (ctypes.c_int32 * 4)(4, 5, 5, 6)  # overlapping pairs; a dummy non-null pointer is never dereferenced on the host
    assert lib.qsv_swap_qubit_pairs(None, 8, pairs, 2, None) == _lib.QSV_EINVAL
    assert lib.qsv_swap_qubit_pairs(ctypes.c_void_p(16), 8, pairs, 2, None) == _lib.QSV_EINVAL
    assert b"disjoint" in lib.qsv_last_error()
    assert lib.qsv_swap_qubit_pairs(ctypes.c_void_p(16), 8, pairs, 0, None) == 0  # nothing to do: no launch


def test_no_cpu_fallback_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from qiskit_sdk_py_b200 import circuits, simulator

    with pytest.raises(_lib.QsvLibraryError, match="no CPU fallback"):
        simulator.QasmSimulatorB200(max_qubits=24).run(circuits.bell())


def test_engine_does_not_import_oracle():
    """Product code must not reach into oracle/ (test infrastructure)."""
    pkg = os.path.join(ROOT, "qiskit_sdk_py_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


def _unitary(rng, k):
    import numpy as np

    z = rng.standard_normal((2 ** k, 2 ** k)) + 1j * rng.standard_normal((2 ** k, 2 ** k))
    q, r = np.linalg.qr(z)
    return q * (np.diag(r) / np.abs(np.diag(r)))


def test_pass_lowering_is_host_only_and_fuses_gate_pairs():
    """qsv_pass_info runs the host lowering of a pass without a GPU: 2-qubit gates that share a qubit are
    merged into 3-qubit blocks (one shared-memory round trip for two gates), gates on the same pair are
    multiplied together, and gates fenced off by an intermediate gate stay separate."""
    import numpy as np

    from qiskit_sdk_py_b200.engine import DeviceState
    from qiskit_sdk_py_b200.lower import GateOp

    rng = np.random.RandomState(1)
    tile = list(range(11))
    dense = lambda a, b: GateOp(_lib.GATE_DENSE2, [a, b], _unitary(rng, 2))
    # chain (4,5),(5,6): one block
    info = DeviceState.pass_info(20, tile, [dense(4, 5), dense(5, 6)])
    assert (info["gates"], info["ops"], info["fused3"]) == (2, 1, 1)
    # same pair twice: one 2-qubit op, no block
    info = DeviceState.pass_info(20, tile, [dense(4, 5), dense(5, 4)])
    assert (info["ops"], info["fused3"]) == (1, 0)
    # (4,5) (6,7) (5,6): (5,6) cannot be hoisted over (6,7) to join (4,5), but it can join (6,7)
    info = DeviceState.pass_info(20, tile, [dense(4, 5), dense(6, 7), dense(5, 6)])
    assert (info["ops"], info["fused3"]) == (2, 1)
    # a third gate inside the block's bits is absorbed
    info = DeviceState.pass_info(20, tile, [dense(4, 5), dense(5, 6), dense(4, 6)])
    assert (info["ops"], info["fused3"]) == (1, 1)
    # disjoint gates are never merged (that would double the FP64 work)
    info = DeviceState.pass_info(20, tile, [dense(4, 5), dense(6, 7), dense(8, 9)])
    assert (info["ops"], info["fused3"], info["stages"]) == (3, 0, 1)
    # copy pieces: 256 bytes (16 amplitudes) when the tile holds qubits 0..3, smaller otherwise
    assert info["piece_log2_load"] == 4 and info["piece_log2_store"] == 4 and info["threads"] == 256
    assert DeviceState.pass_info(20, [0, 1, 2, 9, 10, 11, 12, 13], [])["piece_log2_load"] == 3
    assert DeviceState.pass_info(20, [5, 6, 7, 8], [])["piece_log2_load"] == 0
    # remapped write-back keeps only the unpermuted leading bits in one piece
    info = DeviceState.pass_info(20, tile, [], dest=[0, 1, 3, 2] + list(range(4, 11)))
    assert info["piece_log2_store"] == 2
    # argument errors are the same as qsv_apply_pass's
    with pytest.raises(_lib.QsvLibraryError, match="not a tile qubit"):
        DeviceState.pass_info(20, tile, [dense(4, 15)])


GROUP_THREADS = 3 * 256 + 3 * 32  # three groups of 8 compute warps + one TMA producer warp each


def test_pass_kernel_variant_selection():
    """Host-only: which passes run on the three-group double-buffered variant of the pass kernel (864 threads,
    six tile buffers) and which on the single-buffer variant (one warp per 2^8 amplitudes of the tile)."""
    from qiskit_sdk_py_b200.engine import DeviceState

    tile = list(range(11))
    big = DeviceState.pass_info(22, tile, [])
    assert big["threads"] == GROUP_THREADS and big["smem_bytes"] > 6 * 2048 * 16
    assert DeviceState.pass_info(21, tile, [])["threads"] == 256                       # too few tiles
    assert DeviceState.pass_info(24, list(range(12)), [])["threads"] == 512            # 12-qubit tile
    assert DeviceState.pass_info(24, list(range(10)), [])["threads"] == 128            # 10-qubit tile
    assert DeviceState.pass_info(24, [0, 1] + list(range(5, 14)), [])["threads"] == 256  # 64-byte pieces
    assert DeviceState.pass_info(24, [0, 1, 2] + list(range(5, 13)), [])["threads"] == GROUP_THREADS  # 128-byte pieces
    remap = DeviceState.pass_info(24, tile, [], dest=[1, 0] + list(range(2, 11)))
    assert remap["threads"] == 256                                                     # permuted write-back
    lib = _lib.load()
    assert lib.qsv_debug_set_flags(16) == 0
    try:
        assert DeviceState.pass_info(22, tile, [])["threads"] == 256
        # the bits that switch loads / stores / gates off exist only in the debug library
        assert lib.qsv_debug_set_flags(1) == _lib.QSV_EINVAL
        assert DeviceState.pass_info(22, tile, [])["threads"] == 256
    finally:
        assert lib.qsv_debug_set_flags(0) == 0
    assert DeviceState.pass_info(22, tile, [])["threads"] == GROUP_THREADS


def test_environment_cannot_change_results(monkeypatch):
    """the production library must ignore QSV_DEBUG_FLAGS (flags 1/2/4 would silently corrupt results)."""
    from qiskit_sdk_py_b200.engine import DeviceState

    monkeypatch.setenv("QSV_DEBUG_FLAGS", "23")
    assert DeviceState.pass_info(22, list(range(11)), [])["threads"] == GROUP_THREADS


def test_quantum_volume_pass_statistics():
    """QuantumVolume(33) through planner + lowering, all on the host: about three quarters of the SU(4) gates
    end up in fused blocks and a pass needs fewer than 1.5 stages on average."""
    from qiskit_sdk_py_b200 import circuits, lower, planner
    from qiskit_sdk_py_b200.engine import DeviceState

    n = 33
    qobj = circuits.quantum_volume(n, measure_final=False, shots=1)
    ops = [lower.lower_gate(i["name"], i["qubits"], i.get("params", [])) for i in qobj["experiments"][0]["instructions"]]
    passes = planner.plan_passes(ops, n)
    tot = {"gates": 0, "ops": 0, "fused3": 0, "stages": 0}
    for p in passes:
        info = DeviceState.pass_info(n, p.tile_qubits, p.ops)
        for key in tot:
            tot[key] += info[key]
        # every pass fits the three-group double-buffered variant: six tile buffers + the pass's tables
        assert info["threads"] == GROUP_THREADS and info["smem_bytes"] <= 227 * 1024 - 8400
    assert tot["gates"] == 528
    assert tot["gates"] - tot["ops"] >= 0.35 * tot["gates"]
    assert tot["stages"] <= 1.5 * len(passes)


def test_pass_lowering_fuzz():
    """Random tiles / gate lists / write-back permutations through the host-only lowering (qsv_pass_info):
    no crash, and the invariants of the lowering hold (tools/fuzz_lowering.py is the long version, also run
    under AddressSanitizer)."""
    import subprocess
    import sys

    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_lowering.py"), "7", "400"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "done bad= 0" in out.stdout
