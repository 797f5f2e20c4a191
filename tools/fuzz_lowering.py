#!/usr/bin/env python
"""Fuzz of the HOST-side pass lowering (qsv_pass_info: no GPU touched): random tiles, gate lists and write-back
permutations; checks that nothing crashes and the invariants of the lowering hold.
    python tools/fuzz_lowering.py [seed] [cases]
Under AddressSanitizer (host code of libqsv): build the three .cu files with
    nvcc -O1 -g -std=c++17 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-fsanitize=address -I include -I qiskit_sdk_py_b200/csrc -c ...
link them with `-shared -Xcompiler -fsanitize=address` into /tmp/asan/libqsv.so and run
    QSV_FUZZ_LIB=/tmp/asan/libqsv.so LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python tools/fuzz_lowering.py
(round 1: 9000 cases clean, with and without ASan)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from qiskit_sdk_py_b200 import _lib
if os.environ.get("QSV_FUZZ_LIB"):
    _lib.LIB_PATH = os.environ["QSV_FUZZ_LIB"]
from qiskit_sdk_py_b200.engine import DeviceState
from qiskit_sdk_py_b200.lower import GateOp
rng=np.random.RandomState(int(sys.argv[1]) if len(sys.argv)>1 else 0)
def ru(k):
    z=rng.standard_normal((2**k,2**k))+1j*rng.standard_normal((2**k,2**k)); q,r=np.linalg.qr(z); return q
bad=0
for it in range(int(sys.argv[2]) if len(sys.argv)>2 else 2000):
    n=int(rng.randint(1,36))
    b=int(rng.randint(1,min(n,12)+1))
    tile=sorted(int(x) for x in rng.choice(n,b,replace=False))
    if rng.rand()<0.5 and b>=4 and n>=b:
        low=list(range(min(4,b))); rest=[q for q in range(n) if q not in low]
        tile=low+sorted(int(x) for x in rng.choice(rest,b-len(low),replace=False)) if len(rest)>=b-len(low) else tile
    ng=int(rng.randint(0,49))
    ops=[]
    for _ in range(ng):
        kind=rng.randint(5)
        if b==1 and kind in (1,3,4): kind=0
        if kind==0: ops.append(GateOp(_lib.GATE_DENSE1,[int(rng.choice(tile))],ru(1)))
        elif kind==1:
            a,c=(int(x) for x in rng.choice(tile,2,replace=False)); ops.append(GateOp(_lib.GATE_DENSE2,[a,c],ru(2)))
        elif kind==2: ops.append(GateOp(_lib.GATE_DIAG1,[int(rng.choice(tile))],np.exp(1j*rng.uniform(0,6,2))))
        elif kind==3:
            a,c=(int(x) for x in rng.choice(tile,2,replace=False)); ops.append(GateOp(_lib.GATE_DIAG2,[a,c],np.exp(1j*rng.uniform(0,6,4))))
        else:
            a,c=(int(x) for x in rng.choice(tile,2,replace=False)); ops.append(GateOp(_lib.GATE_CX,[a,c],None))
    dest=None
    if rng.rand()<0.3:
        dest=list(tile); rng.shuffle(dest); dest=[int(x) for x in dest]
    try:
        info=DeviceState.pass_info(n,tile,ops,dest=dest)
    except _lib.QsvLibraryError as e:
        msg=str(e)
        if "payload too large" in msg or "grid too large" in msg: continue
        print("ERR",n,tile,len(ops),msg); bad+=1; continue
    assert info["gates"]==len(ops) and info["ops"]<=max(1,len(ops)) and info["fused3"]<=info["ops"], info
    assert info["threads"] in (32,64,128,256,512,864), info
    assert info["smem_bytes"]<=232448, info
    assert info["stages"]<=max(1,info["ops"]), info
print("done bad=",bad)
