"""(De)serialisation of golden fixtures: a Qobj dict + the reference's outputs, as JSON.

Complex arrays are stored as hex-encoded float64 bytes so the values round-trip bit-exactly.
"""
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _enc(obj):
    if isinstance(obj, np.ndarray):
        if np.iscomplexobj(obj):
            arr = np.ascontiguousarray(obj, dtype=np.complex128)
            return {"__nd__": "c16", "shape": list(arr.shape), "hex": arr.tobytes().hex()}
        if obj.dtype.kind == "f":
            arr = np.ascontiguousarray(obj, dtype=np.float64)
            return {"__nd__": "f8", "shape": list(arr.shape), "hex": arr.tobytes().hex()}
        return {"__nd__": "i8", "shape": list(obj.shape),
                "hex": np.ascontiguousarray(obj, dtype=np.int64).tobytes().hex()}
    if isinstance(obj, complex):
        return {"__complex__": [obj.real, obj.imag]}
    if isinstance(obj, (np.complexfloating,)):
        return {"__complex__": [float(obj.real), float(obj.imag)]}
    if isinstance(obj, (np.integer,)):
        return int(obj)
    if isinstance(obj, (np.floating,)):
        return float(obj)
    if isinstance(obj, (np.bool_,)):
        return bool(obj)
    if isinstance(obj, dict):
        return {str(k): _enc(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [_enc(v) for v in obj]
    if hasattr(obj, "value") and obj.__class__.__module__.startswith("qiskit"):  # enums (MeasLevel)
        return _enc(obj.value)
    return obj


def _dec(obj):
    if isinstance(obj, dict):
        if "__nd__" in obj:
            dt = {"c16": np.complex128, "f8": np.float64, "i8": np.int64}[obj["__nd__"]]
            return np.frombuffer(bytes.fromhex(obj["hex"]), dtype=dt).reshape(obj["shape"]).copy()
        if "__complex__" in obj:
            return complex(obj["__complex__"][0], obj["__complex__"][1])
        return {k: _dec(v) for k, v in obj.items()}
    if isinstance(obj, list):
        return [_dec(v) for v in obj]
    return obj


def dump_case(name, case):
    path = os.path.join(GOLDEN_DIR, name + ".json")
    with open(path, "w") as fh:
        json.dump(_enc(case), fh, separators=(",", ":"))
    return path


def load_case(name):
    path = os.path.join(GOLDEN_DIR, name + ".json")
    with open(path) as fh:
        return _dec(json.load(fh))


def list_cases():
    """Names of the Qobj fixtures (the expectation-value fixture has its own layout and tests)."""
    return sorted(f[:-5] for f in os.listdir(GOLDEN_DIR) if f.endswith(".json") and not f.startswith("expval_"))
