"""Multi-threaded CPU baseline: the oracle's solve loop compiled with numba.

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/mna_oracle.py for the rules).  Same algorithm as the
reference's solver -- for every port and every frequency one LAPACK `zgelsd` least-squares solve of
`As[1:,1:,f] x = e_{N+port}` through numba's `np.linalg.lstsq`, then the power-wave S formula
(src/heavi/solvers/spfn.py:145-176) -- threaded over frequency with `prange` exactly like the
reference threads it, so that `bench.py --impl reference` / `cpu_baseline` time what a user of the
reference gets on the same host cores.  Assembly stays the numpy code of `mna_oracle.assemble`,
as in the reference (Python loop over components, numpy per component).
"""
from __future__ import annotations

import os

import numpy as np

os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_heavi_oracle")

from numba import get_num_threads, njit, prange  # noqa: E402

from . import mna_oracle as orc  # noqa: E402


@njit(parallel=True, fastmath=True, cache=True)
def _solve_all(As, Zs, idx, N, M):
    nports = idx.shape[0]
    nF = As.shape[2]
    NM = N + M
    S = np.zeros((nports, nports, nF), dtype=np.complex128)
    for fi in prange(nF):
        A = np.ascontiguousarray(As[1:, 1:, fi])
        for i in range(nports):
            rhs = np.zeros((NM - 1,), dtype=np.complex128)
            rhs[N + idx[i, 0] - 1] = 1.0
            sol = np.linalg.lstsq(A, rhs)[0]
            x = np.zeros((NM,), dtype=np.complex128)
            x[1:] = sol
            Zi = Zs[idx[i, 0], fi]
            den = (x[idx[i, 1]] - x[idx[i, 3]]) - Zi * (x[idx[i, 1]] - x[idx[i, 2]]) / Zi
            for j in range(nports):
                Zj = Zs[idx[j, 0], fi]
                v1 = x[idx[j, 1]]
                num = (v1 - x[idx[j, 3]]) + np.conj(Zj) * (v1 - x[idx[j, 2]]) / Zj
                S[j, i, fi] = (num / den) * np.sqrt(np.abs(Zi.real)) / np.sqrt(np.abs(Zj.real))
    return S


def threads() -> int:
    return int(get_num_threads())


def run_sp(records, N: int, M: int, f, chunk: int | None = None) -> np.ndarray:
    """Reference-equivalent CPU run_sp on all host threads -> S[P,P,nF]."""
    f = np.asarray(f, dtype=np.float64)
    nF = f.shape[0]
    if chunk is None:
        chunk = max(1, min(nF, int(2e8 // (16 * (N + M) ** 2)) or 1))
    parts = []
    for lo in range(0, nF, chunk):
        fc = f[lo:lo + chunk]
        As = orc.assemble(records, N + M, fc)
        idx, Zs = orc.port_tables(records, M, fc)
        parts.append(_solve_all(As, Zs, idx.astype(np.int64), N, M))
    return np.ascontiguousarray(np.concatenate(parts, axis=2))
