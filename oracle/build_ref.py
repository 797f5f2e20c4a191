#!/usr/bin/env python
"""Compile the reference's own native code for the (next-row) Pauli expectation path into oracle/_ref/.

The only native source near the hot path is qiskit/quantum_info/states/cython/exp_value.pyx.
It is compiled FROM WHERE IT LIES under /root/reference (nothing is copied into the repo): Cython
translates it to C++ in a temporary directory, g++ builds oracle/_ref/exp_value.so.  oracle/_ref/ is
git-ignored.  `cpow=True` is needed because Cython 3 changed the typing of `2 ** x` (exp_value.pyx:58,120).
Used only by tests/golden/make_golden.py (to pin the oracle's expval restatement) -- TEST INFRASTRUCTURE.
"""
import os
import subprocess
import sys
import sysconfig
import tempfile

REF = os.environ.get("QISKIT_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
PYX = os.path.join(REF, "qiskit", "quantum_info", "states", "cython", "exp_value.pyx")


def build():
    if not os.path.exists(PYX):
        return None
    os.makedirs(OUT_DIR, exist_ok=True)
    out = os.path.join(OUT_DIR, "exp_value" + sysconfig.get_config_var("EXT_SUFFIX"))
    if os.path.exists(out) and os.path.getmtime(out) > os.path.getmtime(PYX):
        return out
    import numpy as np

    with tempfile.TemporaryDirectory() as tmp:
        cpp = os.path.join(tmp, "exp_value.cpp")
        subprocess.run([sys.executable, "-m", "cython", "--cplus", "-3", "-X", "cpow=True", PYX, "-o", cpp],
                       check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        subprocess.run(["g++", "-O2", "-funroll-loops", "-std=c++11", "-shared", "-fPIC", "-w",
                        "-I" + sysconfig.get_paths()["include"], "-I" + np.get_include(), cpp, "-o", out], check=True)
    return out


if __name__ == "__main__":
    print(build())
