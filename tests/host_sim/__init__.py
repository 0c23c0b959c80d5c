"""TEST INFRASTRUCTURE: compile the CUDA source of a generated kernel (kernel class 3,
`engine.codegen_source`) for the CPU with g++ and a small CUDA shim, and run it on host arrays.
This checks the generator's algorithm and bookkeeping without a GPU; it is never on a product path."""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_CACHE = os.path.join(tempfile.gettempdir(), "hvb_host_sim")


def build(src: str):
    os.makedirs(_CACHE, exist_ok=True)
    text = '#include "cuda_shim.h"\n' + src + open(os.path.join(HERE, "sim_driver.inc")).read()
    key = hashlib.sha1(text.encode()).hexdigest()[:16]
    so = os.path.join(_CACHE, f"sim_{key}.so")
    if not os.path.exists(so):
        cu = os.path.join(_CACHE, f"sim_{key}.cpp")
        with open(cu, "w") as fh:
            fh.write(text)
        subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-mfma", "-ffp-contract=off", "-Wno-unknown-pragmas",
                        "-Wno-attributes", "-I", HERE, cu, "-o", so + ".tmp"], check=True)
        os.replace(so + ".tmp", so)
    return C.CDLL(so)


def run(gen: dict, P: int, f, consts, prefs, tables, samples, plan):
    """S[nS,P,P,nF] complex128, float32 axis and status word from the generated kernel run on the CPU."""
    lib = build(gen["src"])
    f = np.ascontiguousarray(f, dtype=np.float64)
    samples = np.ascontiguousarray(samples, dtype=np.float64)
    nS, nF = samples.shape[0], f.shape[0]
    S = np.zeros((nS, P, P, nF), dtype=np.complex128)
    f32 = np.zeros(nF, dtype=np.float32)
    status = np.zeros(1, dtype=np.int32)
    keep = [np.ascontiguousarray(consts, dtype=np.complex128), np.ascontiguousarray(prefs, dtype=np.int32),
            np.ascontiguousarray(tables, dtype=np.complex128), np.ascontiguousarray(plan.spl_t, dtype=np.float64),
            np.ascontiguousarray(plan.spl_c, dtype=np.complex128), np.ascontiguousarray(plan.spl_trace, dtype=np.int32),
            np.ascontiguousarray(plan.spl_fill, dtype=np.complex128), np.ascontiguousarray(plan.spl_range, dtype=np.float64),
            np.ascontiguousarray(gen["rows"], dtype=np.int32), np.ascontiguousarray(gen["itab"], dtype=np.int32),
            np.ascontiguousarray(gen["dtab"], dtype=np.float64)]

    def p(a):
        return C.c_void_p(a.ctypes.data) if a.size else None

    lib.sim_run.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int] + [C.c_void_p] * 14
    lib.sim_run(p(f), nF, p(samples), nS, samples.shape[1], *[p(a) for a in keep], p(S), p(f32), p(status))
    return S, f32, int(status[0])
