"""Shared test helpers: golden loading, circuit registry, tolerance + adjudication."""
from __future__ import annotations

import os

import numpy as np

import circuits

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

# north-star tolerance: |S - S_ref| <= 1e-12 + 1e-10 |S_ref| in complex128
ATOL, RTOL = 1e-12, 1e-10


def golden(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def registry(hv):
    """name -> zero-argument builder returning the model (same construction the goldens used)."""
    reg = {
        "c1_readme": lambda: circuits.build_c1(hv)[0],
        "c3_touchstone": lambda: circuits.build_c3(hv, nF=16)[0],
        "c5_ladder": lambda: circuits.build_c5(hv, nF=16)[0],
        "ka_divider": lambda: circuits.build_divider(hv)[0],
        "ka_line": lambda: circuits.build_matched_line(hv)[0],
    }
    for band in ("bandpass", "lowpass", "highpass", "bandstop"):
        reg["c2_" + band] = (lambda b=band: circuits.build_c2(hv, b)[0])
    for seed in range(48):
        reg["rnd_%02d" % seed] = (lambda s=seed: circuits.build_random(hv, s)[0])
    return reg


def dense_y(g):
    NM = int(g["N"]) + int(g["M"])
    Y = np.zeros((NM, NM, len(g["y_f"])), dtype=np.complex128)
    Y[g["y_rows"], g["y_cols"], g["y_fidx"]] = g["y_vals"]
    return Y


def max_ulps(a, b):
    d = np.abs(a - b)
    s = np.maximum(np.abs(a), np.abs(b))
    s[s == 0] = 1
    return float((d / s).max() / np.finfo(np.float64).eps) if d.size else 0.0


def parity_report(S, S_ref, S_ext):
    """Pass fraction on the plain tolerance, and the adjudication of every failing point: the
    candidate must be at least as close to the extended-precision solve as the reference is
    (SURVEY.md App. E -- the reference's SVD least squares is the noisy side)."""
    tol = ATOL + RTOL * np.abs(S_ref)
    bad = np.abs(S - S_ref) > tol
    bad_pts = np.any(bad, axis=tuple(range(S.ndim - 1)))
    frac = 1.0 - bad_pts.mean() if bad_pts.size else 1.0
    adjudicated = True
    if bad.any():
        adjudicated = bool(np.all(np.abs(S - S_ext)[bad] <= np.abs(S_ref - S_ext)[bad] + 1e-15))
    return frac, adjudicated, float(np.abs(S - S_ext).max())
