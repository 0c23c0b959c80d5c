"""CPU oracle for heavi's frequency-sweep S-parameter path (numpy restatement).

TEST INFRASTRUCTURE ONLY.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import this module, and only as the checker or the
timed CPU baseline.  The product package `heavi_b200` never imports it: the product path is the
CUDA library and fails loudly when that library is missing.

What is restated (file:line are relative to /root/reference/src/heavi):

* stamp assembly, one `+=` of a dense (k,k,nF) block per component into a dense
  (N+M, N+M, nF) complex128 tensor                       network.py:392-394, base/*.py
* port table `indices[P,4]` and `impedance_vector[M,nF]`  network.py:397-404
* per port, per frequency least-squares solve of `As[1:,1:,f] x = e_{N+port}` followed by the
  power-wave S formula                                    solvers/spfn.py:145-176

The reference's solve is numba `np.linalg.lstsq` = LAPACK `zgelsd` with rcond = machine eps
(third party: numba 0.61.0 / SciPy LAPACK, pinned in uv.lock:346,560).  `method="lstsq"` calls the
same LAPACK driver through numpy; `method="lu"` is partial-pivot LU (`zgesv`), which is what the
GPU path computes and which SURVEY.md App. E shows agrees with the reference to the reference's
own SVD rounding noise.  `solve_extended` is the adjudicator: the same linear systems solved in
x87 80-bit arithmetic.

Pinning: `oracle/check_oracle_vs_ref.py` runs the real reference (through `oracle/ref_shim.py`)
and this module on the same circuits; `oracle/make_golden.py` writes the reference's own outputs
to `tests/golden/*.npz`, against which `tests/test_oracle_golden.py` checks this module on every
machine.  The reference ships no tests or golden vectors of its own.

A circuit is given as neutral *records* (plain dicts), so that the oracle does not depend on the
product package:

    {"kind": "port",       "ids": [sig, int, gnd, src], "Z0": p, "port": k}     base/port.py:23-42
    {"kind": "resistor",   "ids": [a, b], "G": p}                                base/resistor.py:51-58
    {"kind": "admittance", "ids": [a, b], "Y": p}                                base/admittance.py:35-42
    {"kind": "impedance",  "ids": [a, b], "Z": p}                                base/impedance.py:25-32
    {"kind": "capacitor",  "ids": [a, b], "C": p}                                base/capacitor.py:27-34
    {"kind": "inductor",   "ids": [a, b], "L": p}                                base/inductor.py:28-35
    {"kind": "tline",      "ids": [a, b, g], "Z0": p, "beta": p, "length": x}    base/tl.py:25-67
    {"kind": "nport_s",    "ids": [n1..nP, g], "S": [[p]], "Z0": x}              base/nport.py:24-45

where `p` is a callable of the float64 frequency array (the value semantics of
numeric.py:127-129: returns an array of len(f) or a 0-d value) and `x` a plain number.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_TWO = np.array([[1.0, -1.0], [-1.0, 1.0]])


# --------------------------------------------------------------------------------------------
# stamp blocks, one function per component class
# --------------------------------------------------------------------------------------------

def _block_port(rec, f):
    # base/port.py:29-39 and base/numeric.py:4-10: float64 block, g = 1/Z0(f)
    t_cond = np.array([[1, -1, 0, 0], [-1, 1, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0]], dtype=np.float64)
    t_src = np.array([[0, 0, 0, 0], [0, 0, 0, 1], [0, 0, 0, -1], [0, 1, -1, 0]], dtype=np.float64)
    g = 1 / rec["Z0"](f)
    g = np.broadcast_to(np.asarray(g, dtype=np.float64), f.shape)
    return t_cond[:, :, None] * g[None, None, :] + t_src[:, :, None]


def _two_terminal(y, f):
    y = np.broadcast_to(np.asarray(y, dtype=np.complex128), f.shape)
    return _TWO[:, :, None] * y[None, None, :]


def _block_resistor(rec, f):      # base/resistor.py:54-56
    return rec["G"](f) * _TWO[:, :, np.newaxis] * np.ones((1, 1, f.shape[0]))


def _block_admittance(rec, f):    # base/admittance.py:38-40
    return _two_terminal(rec["Y"](f), f)


def _block_impedance(rec, f):     # base/impedance.py:28-30
    return _two_terminal(1 / rec["Z"](f), f)


def _block_capacitor(rec, f):     # base/capacitor.py:30-32
    return _two_terminal(1j * 2 * np.pi * f * rec["C"](f), f)


def _block_inductor(rec, f):      # base/inductor.py:31-33
    return _two_terminal(1 / (1j * 2 * np.pi * f * rec["L"](f)), f)


def _pad_indefinite(Y):
    """(k,k,nF) port admittance -> (k+1,k+1,nF) indefinite block with the reference row/column
    (base/tl.py:53-59, base/nport.py:33-38)."""
    k, _, nF = Y.shape
    Y2 = np.zeros((k + 1, k + 1, nF), dtype=np.complex128)
    Y2[:k, :k, :] = Y
    for i in range(k):
        Y2[i, k, :] = -np.sum(Y[i, :, :], axis=0)
        Y2[k, i, :] = -np.sum(Y[:, i, :], axis=0)
        Y2[k, k, :] += np.sum(Y[i, :, :], axis=0)
    return Y2


def _block_tline(rec, f):         # base/tl.py:30-60
    beta, Z0, length = rec["beta"], rec["Z0"], rec["length"]
    th = 1j * beta(f) * length
    a11 = np.cosh(th)
    a12 = Z0(f) * np.sinh(th)
    a21 = 1 / Z0(f) * np.sinh(th)
    a22 = np.cosh(th)
    y11 = a22 / a12
    y12 = -((a11 * a22) - (a12 * a21)) / a12
    y21 = -1 / a12
    y22 = a11 / a12
    nF = f.shape[0]
    Y = np.empty((2, 2, nF), dtype=np.complex128)
    Y[0, 0], Y[0, 1], Y[1, 0], Y[1, 1] = y11, y12, y21, y22
    return _pad_indefinite(Y)


def _block_nport_s(rec, f):       # base/nport.py:28-39
    funcs = rec["S"]
    P = len(funcs)
    nF = f.shape[0]
    S = np.empty((P, P, nF), dtype=np.complex128)
    for i in range(P):
        for j in range(P):
            S[i, j, :] = funcs[i][j](f)
    eye = np.eye(P)[:, :, None]
    inv = np.stack([np.linalg.inv((eye + S)[:, :, m]) for m in range(nF)], axis=2)
    Y = (1 / rec["Z0"]) * np.einsum("ijk,jlk->ilk", (eye - S), inv)
    return _pad_indefinite(Y)


_BLOCKS = {
    "port": _block_port,
    "resistor": _block_resistor,
    "admittance": _block_admittance,
    "impedance": _block_impedance,
    "capacitor": _block_capacitor,
    "inductor": _block_inductor,
    "tline": _block_tline,
    "nport_s": _block_nport_s,
}


def assemble(records, NM: int, f: np.ndarray) -> np.ndarray:
    """Dense MNA tensor (NM, NM, nF), network.py:392-394.

    numpy fancy-index `+=` is read-add-write per component, so a component whose `ids` repeat an
    index keeps only the last (row-major) write for that cell -- exactly what the reference does.
    """
    f = np.asarray(f, dtype=np.float64)
    Y = np.zeros((NM, NM, f.shape[0]), dtype=np.complex128)
    for rec in records:
        ids = np.asarray(rec["ids"])
        row, col = np.ix_(ids, ids)
        Y[row, col, :] += _BLOCKS[rec["kind"]](rec, f)
    return Y


def port_tables(records, M: int, f: np.ndarray):
    """`indices[P,4]` int32 and `impedance_vector[M,nF]` complex128, network.py:397-404."""
    f = np.asarray(f, dtype=np.float64)
    Zs = np.zeros((M, f.shape[0]), dtype=np.complex128)
    rows = []
    ports = sorted((r for r in records if r["kind"] == "port"), key=lambda r: r["port"])
    for rec in ports:
        k = rec["port"]
        sig, internal, gnd, _src = rec["ids"]
        rows.append((k - 1, sig, internal, gnd))
        Zs[k - 1, :] = rec["Z0"](f)
    return np.array(rows).astype(np.int32), Zs


# --------------------------------------------------------------------------------------------
# solve + S conversion, solvers/spfn.py:129-179
# --------------------------------------------------------------------------------------------

def _s_from_x(x, Zs_f, idx, i, out):
    """spfn.py:159-176 for one excitation i; x is the full solution vector incl. ground slot 0."""
    Z_in = Zs_f[idx[i, 0]]
    Vi1, Vi2, Vi3 = x[idx[i, 1]], x[idx[i, 3]], x[idx[i, 2]]
    for j in range(idx.shape[0]):
        Z_out = Zs_f[idx[j, 0]]
        Vo1, Vo2, Vo3 = x[idx[j, 1]], x[idx[j, 3]], x[idx[j, 2]]
        num = (Vo1 - Vo2) + np.conj(Z_out) * (Vo1 - Vo3) / Z_out
        den = (Vi1 - Vi2) - Z_in * (Vi1 - Vi3) / Z_in
        out[j, i] = (num / den) * np.sqrt(np.abs(np.real(Z_in))) / np.sqrt(np.abs(np.real(Z_out)))


def _solve_block(As, Zs, idx, N, M, method, lo, hi, S):
    NM = N + M
    P = idx.shape[0]
    for fi in range(lo, hi):
        A = np.ascontiguousarray(As[1:, 1:, fi])
        if method == "lu":
            B = np.zeros((NM - 1, P), dtype=np.complex128)
            for i in range(P):
                B[N + idx[i, 0] - 1, i] = 1
            X = np.linalg.solve(A, B)
        for i in range(P):
            x = np.zeros(NM, dtype=np.complex128)
            if method == "lstsq":
                z = np.zeros(NM, dtype=np.complex128)
                z[N + idx[i, 0]] = 1
                x[1:] = np.linalg.lstsq(A, z[1:], rcond=-1)[0]
            else:
                x[1:] = X[:, i]
            _s_from_x(x, Zs[:, fi], idx, i, S[:, :, fi])


def solve_sp(As, Zs, idx, N: int, M: int, method: str = "lstsq", threads: int = 1) -> np.ndarray:
    """S[P,P,nF] from the dense tensor; one solve per port per frequency like spfn.py:145-156."""
    if method not in ("lstsq", "lu"):
        raise ValueError(method)
    nF = As.shape[2]
    P = idx.shape[0]
    S = np.zeros((P, P, nF), dtype=np.complex128)
    if threads <= 1 or nF < 4 * threads:
        _solve_block(As, Zs, idx, N, M, method, 0, nF, S)
        return S
    edges = np.linspace(0, nF, threads + 1).astype(int)
    with ThreadPoolExecutor(threads) as pool:
        futs = [pool.submit(_solve_block, As, Zs, idx, N, M, method, edges[t], edges[t + 1], S)
                for t in range(threads)]
        for fu in futs:
            fu.result()
    return S


def solve_sp_lu_batched(As, Zs, idx, N: int, M: int) -> np.ndarray:
    """Vectorised LU variant (all frequencies in one LAPACK batch) for large parity cases.
    Same arithmetic as method="lu": zgesv on As[1:,1:,f] with the P unit right-hand sides."""
    NM = N + M
    P = idx.shape[0]
    nF = As.shape[2]
    A = np.ascontiguousarray(np.moveaxis(As[1:, 1:, :], 2, 0))
    B = np.zeros((nF, NM - 1, P), dtype=np.complex128)
    for i in range(P):
        B[:, N + idx[i, 0] - 1, i] = 1
    X = np.linalg.solve(A, B)                       # (nF, NM-1, P)
    Xf = np.zeros((nF, NM, P), dtype=np.complex128)
    Xf[:, 1:, :] = X
    return s_from_solution(Xf, Zs, idx)


def s_from_solution(Xf, Zs, idx) -> np.ndarray:
    """Vectorised spfn.py:159-176.  Xf[(nF, NM, P)] holds the solution incl. the ground slot."""
    P = idx.shape[0]
    nF = Xf.shape[0]
    S = np.zeros((P, P, nF), dtype=Xf.dtype)
    for i in range(P):
        Z_in = Zs[idx[i, 0], :]
        Vi1, Vi2, Vi3 = Xf[:, idx[i, 1], i], Xf[:, idx[i, 3], i], Xf[:, idx[i, 2], i]
        den = (Vi1 - Vi2) - Z_in * (Vi1 - Vi3) / Z_in
        for j in range(P):
            Z_out = Zs[idx[j, 0], :]
            Vo1, Vo2, Vo3 = Xf[:, idx[j, 1], i], Xf[:, idx[j, 3], i], Xf[:, idx[j, 2], i]
            num = (Vo1 - Vo2) + np.conj(Z_out) * (Vo1 - Vo3) / Z_out
            S[j, i, :] = (num / den) * np.sqrt(np.abs(np.real(Z_in))) / np.sqrt(np.abs(np.real(Z_out)))
    return S


def run_sp(records, N: int, M: int, f, method: str = "lstsq", threads: int = 1,
           chunk: int | None = None) -> np.ndarray:
    """network.py:380-421 end to end on neutral records -> S[P,P,nF] (complex128, C order).

    `chunk` bounds the dense tensor like a user of the reference would have to (SURVEY.md
    finding 2); frequency points are independent so chunking is exact.
    """
    f = np.asarray(f, dtype=np.float64)
    nF = f.shape[0]
    if chunk is None:
        chunk = max(1, min(nF, int(2e8 // (16 * (N + M) ** 2)) or 1))
    parts = []
    for lo in range(0, nF, chunk):
        fc = f[lo:lo + chunk]
        As = assemble(records, N + M, fc)
        idx, Zs = port_tables(records, M, fc)
        if method == "lu_batched":
            parts.append(solve_sp_lu_batched(As, Zs, idx, N, M))
        else:
            parts.append(solve_sp(As, Zs, idx, N, M, method=method, threads=threads))
    return np.ascontiguousarray(np.concatenate(parts, axis=2))


# --------------------------------------------------------------------------------------------
# extended-precision adjudicator: the same float64 matrices solved in x87
# 80-bit complex arithmetic with partial pivoting, vectorised over frequency.
# --------------------------------------------------------------------------------------------

def solve_extended(As, Zs, idx, N: int, M: int) -> np.ndarray:
    NM = N + M
    P = idx.shape[0]
    nF = As.shape[2]
    n = NM - 1
    A = np.moveaxis(As[1:, 1:, :], 2, 0).astype(np.clongdouble)
    B = np.zeros((nF, n, P), dtype=np.clongdouble)
    for i in range(P):
        B[:, N + idx[i, 0] - 1, i] = 1
    W = np.concatenate([A, B], axis=2)                     # (nF, n, n+P)
    ar = np.arange(nF)
    for k in range(n):
        p = k + np.argmax(np.abs(W[:, k:, k]), axis=1)
        rk = W[ar, k, :].copy()
        W[ar, k, :] = W[ar, p, :]
        W[ar, p, :] = rk
        piv = W[:, k, k]
        if k + 1 < n:
            m = W[:, k + 1:, k] / piv[:, None]
            W[:, k + 1:, k:] -= m[:, :, None] * W[:, k, None, k:]
    X = np.zeros((nF, n, P), dtype=np.clongdouble)
    for k in range(n - 1, -1, -1):
        acc = W[:, k, n:] - np.einsum("fc,fcp->fp", W[:, k, k + 1:n], X[:, k + 1:, :])
        X[:, k, :] = acc / W[:, k, k, None]
    Xf = np.zeros((nF, NM, P), dtype=np.clongdouble)
    Xf[:, 1:, :] = X
    return s_from_solution(Xf, Zs.astype(np.clongdouble), idx)


def host_threads() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
