"""ctypes binding of libqsv.so (include/qsv.h).

There is no CPU fallback: if the shared library is missing or cannot be loaded, every entry point of
the package raises ``QsvLibraryError``.  Build it with ``python __graft_entry__.py`` (or
``make -C qiskit_sdk_py_b200/csrc``).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# QSV_LIB_PATH selects another build of the same library (tools/ use the profiling build libqsv_dbg.so)
LIB_PATH = os.environ.get("QSV_LIB_PATH") or os.path.join(_HERE, "libqsv.so")

QSV_EINVAL = -1
QSV_ERANGE = -2
QSV_ENORM = -3

GATE_DENSE1 = 1
GATE_DENSE2 = 2
GATE_DIAG1 = 3
GATE_DIAG2 = 4
GATE_CX = 5

MAX_TILE_QUBITS = 12
MAX_PASS_GATES = 48
MAX_PASS_MATRIX_DOUBLES = 1536
MAX_DENSE_K = 12
MAX_TAIL_GATES = 192


class QsvLibraryError(RuntimeError):
    """libqsv.so is missing / failed to load, or a call into it failed."""


class QsvGate(ctypes.Structure):
    """``qsv_gate_t`` of include/qsv.h."""

    _fields_ = [
        ("kind", ctypes.c_int32),
        ("q0", ctypes.c_int32),
        ("q1", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("m", ctypes.c_double * 32),
    ]


_c_void_p = ctypes.c_void_p
_c_int = ctypes.c_int
_c_double = ctypes.c_double
_c_size_t = ctypes.c_size_t
_p_int32 = ctypes.POINTER(ctypes.c_int32)
_p_double = ctypes.POINTER(ctypes.c_double)
_p_int64 = ctypes.POINTER(ctypes.c_int64)

# name -> (restype, argtypes); must list every symbol include/qsv.h declares
SIGNATURES = {
    "qsv_version": (_c_int, []),
    "qsv_last_error": (ctypes.c_char_p, []),
    "qsv_launch_count": (ctypes.c_uint64, []),
    "qsv_init_basis0": (_c_int, [_c_void_p, _c_int, _c_double, _c_double, _c_void_p]),
    "qsv_init_identity": (_c_int, [_c_void_p, _c_int, _c_double, _c_double, _c_void_p]),
    "qsv_upload": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_size_t, _c_double, _c_double, _c_void_p]),
    "qsv_scale": (_c_int, [_c_void_p, _c_int, _c_double, _c_double, _c_void_p]),
    "qsv_apply_pass": (_c_int, [_c_void_p, _c_int, _p_int32, _c_int, ctypes.POINTER(QsvGate), _c_int,
                                _c_void_p]),
    "qsv_apply_pass_remap": (_c_int, [_c_void_p, _c_int, _p_int32, _c_int, ctypes.POINTER(QsvGate), _c_int,
                                      _p_int32, _c_void_p]),
    "qsv_apply_pass_tail": (_c_int, [_c_void_p, _c_int, _p_int32, _c_int, ctypes.POINTER(QsvGate), _c_int, _p_int32,
                                     ctypes.POINTER(QsvGate), _c_int, _c_void_p, _c_void_p]),
    "qsv_tail_scratch_bytes": (_c_size_t, []),
    "qsv_apply_pass_sub": (_c_int, [_c_void_p, _c_int, _p_int32, _c_int, ctypes.POINTER(QsvGate), _c_int,
                                    ctypes.POINTER(QsvGate), _c_int, _c_void_p, _p_int32, _c_int, ctypes.c_uint32, _c_int,
                                    _c_void_p]),
    "qsv_pass_info": (_c_int, [_c_int, _p_int32, _c_int, ctypes.POINTER(QsvGate), _c_int, _p_int32, _p_int32]),
    "qsv_pass_ops": (_c_int, [_c_int, _p_int32, _c_int, ctypes.POINTER(QsvGate), _c_int, _p_int32, _p_int32, _c_int]),
    "qsv_apply_matrix": (_c_int, [_c_void_p, _c_int, _p_int32, _c_int, _c_void_p, _c_void_p, _c_void_p]),
    "qsv_reduce_workspace_bytes": (_c_size_t, []),
    "qsv_prob_qubit": (_c_int, [_c_void_p, _c_int, _c_int, _c_void_p, _p_double, _c_void_p]),
    "qsv_collapse": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_double, _c_int, _c_void_p]),
    "qsv_batch_workspace_bytes": (_c_size_t, [_c_int]),
    "qsv_collapse_batch": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_void_p]),
    "qsv_apply_masked": (_c_int, [_c_void_p, _c_int, _c_int, ctypes.POINTER(QsvGate), _c_void_p, _c_void_p, _c_void_p]),
    "qsv_marginal_workspace_bytes": (_c_size_t, [_c_int, _c_int]),
    "qsv_marginal": (_c_int, [_c_void_p, _c_int, _p_int32, _c_int, _c_void_p, _c_void_p, _c_void_p]),
    "qsv_sample_workspace_bytes": (_c_size_t, [_c_int, _c_int]),
    "qsv_sample": (_c_int, [_c_void_p, _c_int, _c_int, _c_void_p, _c_int, _c_void_p, _c_int, _c_void_p,
                            _p_double, _c_void_p]),
    "qsv_cdf_prepare": (_c_int, [_c_void_p, _c_int, _c_int, _c_void_p, _p_double, _c_void_p]),
    "qsv_cdf_classify": (_c_int, [_c_void_p, _c_int, _c_int, _c_double, _c_void_p, _c_void_p]),
    "qsv_cdf_chain": (_c_int, [_c_void_p, _c_int, _c_int, _c_double, _c_void_p, _p_double, _c_void_p]),
    "qsv_cdf_search": (_c_int, [_c_void_p, _c_int, _c_int, _c_double, _c_double, _c_void_p, _c_int, _c_void_p,
                                _c_void_p, _c_void_p]),
    "qsv_chop": (_c_int, [_c_void_p, _c_int, _c_double, _c_void_p]),
    "qsv_download": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_void_p]),
    "qsv_pack_half": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_void_p, _c_size_t, _c_size_t, _c_void_p]),
    "qsv_unpack_half": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_void_p, _c_size_t, _c_size_t, _c_void_p]),
    "qsv_expval_pauli": (_c_int, [_c_void_p, _c_int, ctypes.c_uint64, ctypes.c_uint64, _c_double, _c_double,
                                  _c_void_p, _p_double, _c_void_p]),
    "qsv_pack_bits": (_c_int, [_c_void_p, _c_int, _p_int32, _c_int, ctypes.c_uint32, _c_void_p, _c_size_t, _c_size_t,
                               _c_void_p]),
    "qsv_unpack_bits": (_c_int, [_c_void_p, _c_int, _p_int32, _c_int, ctypes.c_uint32, _c_void_p, _c_size_t,
                                 _c_size_t, _c_void_p]),
    "qsv_norm2": (_c_int, [_c_void_p, _c_int, _c_void_p, _p_double, _c_void_p]),
    "qsv_swap_bits_p2p": (_c_int, [_c_void_p, _c_int, _p_int32, _c_int, ctypes.c_uint32, ctypes.POINTER(_c_void_p), _c_int,
                                  _c_void_p]),
    "qsv_ipc_export": (_c_int, [_c_void_p, _c_void_p, ctypes.POINTER(ctypes.c_uint64)]),
    "qsv_ipc_import": (_c_int, [_c_void_p, ctypes.c_uint64, ctypes.POINTER(_c_void_p)]),
    "qsv_ipc_release": (_c_int, [_c_void_p, ctypes.c_uint64]),
    "qsv_scatter_bits": (_c_int, [_c_void_p, _c_int, _p_int32, ctypes.c_uint64, _c_void_p, _c_int, _c_void_p]),
    "qsv_alloc": (_c_int, [ctypes.POINTER(_c_void_p), _c_size_t]),
    "qsv_free": (_c_int, [_c_void_p]),
    "qsv_copy_d2d": (_c_int, [_c_void_p, _c_void_p, _c_size_t, _c_void_p]),
    "qsv_swap_qubit_pairs": (_c_int, [_c_void_p, _c_int, _p_int32, _c_int, _c_void_p]),
    "qsv_debug_set_flags": (_c_int, [ctypes.c_uint32]),
    "qsv_debug_profile": (_c_int, [ctypes.POINTER(ctypes.c_uint64), _c_size_t]),
}

_lib = None


def load():
    """Load libqsv.so once; raise QsvLibraryError (never fall back) if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise QsvLibraryError(
            "%s not found: the CUDA library is not built (run `python __graft_entry__.py`). "
            "This package has no CPU fallback." % LIB_PATH
        )
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as exc:  # pragma: no cover - depends on the box
        raise QsvLibraryError("cannot load %s: %s" % (LIB_PATH, exc)) from exc
    for name, (restype, argtypes) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as exc:
            raise QsvLibraryError("libqsv.so does not export %s" % name) from exc
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc, what):
    """Turn a non-zero status into an exception carrying qsv_last_error()."""
    if rc == 0:
        return
    msg = load().qsv_last_error().decode("utf-8", "replace")
    if rc == QSV_ENORM:
        # numpy's RandomState.choice raises ValueError("probabilities do not sum to 1")
        raise ValueError("probabilities do not sum to 1")
    raise QsvLibraryError("%s failed (status %d): %s" % (what, rc, msg))


def launch_count():
    return int(load().qsv_launch_count())
