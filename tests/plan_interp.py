"""numpy interpreter of a `DevicePlan` -- an executable specification of what the CUDA kernels do
with the plan arrays (value program -> cell program -> solve -> S).  TEST INFRASTRUCTURE ONLY: it
lets the CPU test-suite (no GPU in the build container) check that the *plan compiler* is right,
so that GPU debugging is confined to the kernels themselves.  The product never imports it.
"""
from __future__ import annotations

import numpy as np

from heavi_b200 import plan as pl

_C0 = 299792458
_TWOPI = np.float64(2 * np.pi)


def _bspline_eval(t, c, k, x):
    """scipy's evaluate_spline / _deBoor_D restated (vectorised over x)."""
    n = len(t) - k - 1
    ell = np.clip(np.searchsorted(t, x, side="right") - 1, k, n - 1)
    h = np.zeros((k + 1,) + x.shape)
    hh = np.zeros_like(h)
    h[0] = 1.0
    for j in range(1, k + 1):
        hh[:j] = h[:j]
        h[0] = 0.0
        for m in range(1, j + 1):
            xb = t[ell + m]
            xa = t[ell + m - j]
            same = xb == xa
            w = np.where(same, 0.0, hh[m - 1] / np.where(same, 1.0, xb - xa))
            h[m - 1] = h[m - 1] + w * (xb - x)
            h[m] = w * (x - xa)
    out = np.zeros(x.shape, dtype=np.complex128)
    for a in range(k + 1):
        out = out + c[ell - k + a] * h[a]
    return out


class Interp:
    def __init__(self, plan: pl.DevicePlan, f, consts, prefs, tables, samples_row):
        self.p = plan
        self.f = np.asarray(f, dtype=np.float64)
        self.consts, self.prefs, self.tables, self.s = consts, prefs, tables, samples_row

    def param(self, pi):
        kind, idx = self.prefs[pi]
        f = self.f
        one = np.ones_like(f)
        if kind == pl.PK_CONST:
            v = self.consts[idx]
            return (v.real if v.imag == 0 else v) * one
        if kind == pl.PK_SAMPLE:
            return self.s[idx] * one
        if kind == pl.PK_SAMPLE_NEG:
            return -self.s[idx] * one
        if kind == pl.PK_SAMPLE_INV:
            return (1 / self.s[idx]) * one
        if kind == pl.PK_TABLE:
            return self.tables[idx]
        if kind == pl.PK_BETA:
            root = self.consts[idx]                      # sqrt(er), taken by the plan compiler
            root = root.real if root.imag == 0 else root
            return _TWOPI * f / _C0 * (root * one)
        if kind == pl.PK_BETA_MUL:
            return 2 * np.pi * f * self.consts[idx].real / _C0
        if kind == pl.PK_SMD:
            sub, a, b, c = (self.consts[idx + i].real for i in range(4))
            w = 2 * 3.141592653589793 * f
            zs = a + 1j * w * b
            zc = 1 / (1j * w * c)
            z = (zs * zc / (zs + zc)) if sub == 0 else (1 / (1 / zs + 1 / zc)) if sub == 1 else (zs + zc)
            return np.nan_to_num(z, posinf=np.inf, neginf=-np.inf)
        if kind == pl.PK_SPLINE:
            t_off, nt, c_off, k = self.p.spl_trace[idx]
            t = self.p.spl_t[t_off:t_off + nt]
            c = self.p.spl_c[c_off:c_off + nt - k - 1]
            v = _bspline_eval(t, c, k, f)
            lo, hi = self.p.spl_range[idx]
            v = np.where(f < lo, self.p.spl_fill[idx, 0], v)
            return np.where(f > hi, self.p.spl_fill[idx, 1], v)
        raise ValueError(kind)

    def values(self):
        p = self.p
        f = self.f
        nvc = len(p.const_vals)
        vals = np.zeros((nvc + p.n_dyn, f.shape[0]), dtype=np.complex128)
        vals[:nvc] = p.const_vals[:, None]
        for op in p.coops:
            code, out, Pn, first = op[0], op[1], op[2], op[3]
            M = np.array([[self.param(first + i * Pn + j) for j in range(Pn)] for i in range(Pn)])  # (P,P,nF)
            if code == pl.COOP_NPORT_S:
                z0 = self.consts[op[4]]
                z0 = z0.real if z0.imag == 0 else z0
                eye = np.eye(Pn)[:, :, None]
                # solve X (I+S) = (I-S)  <=>  (I+S)^T X^T = (I-S)^T, per frequency
                A = np.moveaxis(eye + M, 2, 0).transpose(0, 2, 1)
                B = np.moveaxis(eye - M, 2, 0).transpose(0, 2, 1)
                X = np.linalg.solve(A, B).transpose(0, 2, 1)
                Y = np.moveaxis((1 / z0) * X, 0, 2)
            else:
                Y = M.astype(np.complex128)
            Y2 = np.zeros((Pn + 1, Pn + 1, f.shape[0]), dtype=np.complex128)
            Y2[:Pn, :Pn] = Y
            for i in range(Pn):
                Y2[i, Pn] = -np.sum(Y[i, :, :], axis=0)
                Y2[Pn, i] = -np.sum(Y[:, i, :], axis=0)
                Y2[Pn, Pn] += np.sum(Y[i, :, :], axis=0)
            vals[out:out + (Pn + 1) ** 2] = Y2.reshape((Pn + 1) ** 2, -1)
        for code, out, arg in zip(p.fop_code, p.fop_out, p.fop_arg):
            v = self.s[int(arg)] * np.ones_like(f) if code & 4 else arg * np.ones_like(f)
            op = code & 3
            if op == pl.OP_COPY:
                vals[nvc + out] = v
            elif op == pl.OP_RECIP:
                vals[nvc + out] = 1 / v
            elif op == pl.OP_CAP:
                vals[nvc + out] = 1j * 2 * np.pi * f * v
            else:
                vals[nvc + out] = 1 / (1j * 2 * np.pi * f * v)
        for op in p.ops:
            code, out = op[0], op[1]
            if code == pl.OP_COPY:
                vals[out] = self.param(op[2])
            elif code == pl.OP_RECIP:
                vals[out] = 1 / self.param(op[2])
            elif code == pl.OP_CAP:
                vals[out] = 1j * 2 * np.pi * f * self.param(op[2])
            elif code == pl.OP_IND:
                vals[out] = 1 / (1j * 2 * np.pi * f * self.param(op[2]))
            elif code == pl.OP_TLINE:
                Z0 = self.param(op[2])
                beta = self.param(op[3])
                length = np.real(self.param(op[4]))
                th = 1j * beta * length
                a11 = np.cosh(th)
                a12 = Z0 * np.sinh(th)
                a21 = 1 / Z0 * np.sinh(th)
                a22 = np.cosh(th)
                y11 = a22 / a12
                y12 = -((a11 * a22) - (a12 * a21)) / a12
                y21 = -1 / a12
                y22 = a11 / a12
                s0, s1 = y11 + y12, y21 + y22
                vals[out + 0], vals[out + 1], vals[out + 2], vals[out + 3] = y11, y12, y21, y22
                vals[out + 4], vals[out + 5] = -s0, -s1
                vals[out + 6], vals[out + 7] = -(y11 + y21), -(y12 + y22)
                vals[out + 8] = (0 + s0) + s1
            else:
                raise ValueError(code)
        return vals

    def matrix(self):
        """Dense [n, ncols, nF] system assembled from the cell program."""
        p = self.p
        vals = self.values()
        W = np.zeros((p.n, p.ncols, self.f.shape[0]), dtype=np.complex128)
        for r in range(p.n):
            for c in range(p.ncols):
                lo, hi = p.cell_ptr[r * p.ncols + c], p.cell_ptr[r * p.ncols + c + 1]
                for e in p.cell_ent[lo:hi]:
                    v = vals[e >> 1]
                    W[r, c] = W[r, c] - v if (e & 1) else W[r, c] + v
        return W

    def solve(self):
        """S[P,P,nF] the way the device epilogue forms it (solvers/spfn.py:159-176)."""
        p = self.p
        W = np.moveaxis(self.matrix(), 2, 0)
        X = np.linalg.solve(W[:, :, :p.n], W[:, :, p.n:])             # (nF, n, P)
        nF = self.f.shape[0]
        Z = [self.param(p.port_zpref[q]) for q in range(p.P)]

        def volt(q, which, i):
            r = p.port_rows[q, which]
            if r == pl.ROW_GROUND:
                return np.zeros(nF, dtype=np.complex128)
            if r == pl.ROW_VIRTUAL_INT:
                return volt(q, 2, i) + (1.0 if q == i else 0.0)
            return X[:, r, i]

        if p.epi_simple and getattr(self, "use_closed_form", False):
            # the closed-form epilogue kernel A uses: S_ji = (2 V_sig_j - delta_ij) sqrt|Zi|/sqrt|Zj|
            S = np.zeros((p.P, p.P, nF), dtype=np.complex128)
            for r in range(p.n):
                j = p.epi_port_of_row[r]
                if j < 0:
                    continue
                for i in range(p.P):
                    S[j, i] = (2 * X[:, r, i] - (1.0 if i == j else 0.0)) * p.epi_scale[j, i]
            return S
        S = np.zeros((p.P, p.P, nF), dtype=np.complex128)
        for i in range(p.P):
            Vi1, Vi3, Vi2 = volt(i, 0, i), volt(i, 1, i), volt(i, 2, i)
            den = (Vi1 - Vi2) - Z[i] * (Vi1 - Vi3) / Z[i]
            for j in range(p.P):
                Vo1, Vo3, Vo2 = volt(j, 0, i), volt(j, 1, i), volt(j, 2, i)
                num = (Vo1 - Vo2) + np.conj(Z[j]) * (Vo1 - Vo3) / Z[j]
                S[j, i] = (num / den) * np.sqrt(np.abs(np.real(Z[i]))) / np.sqrt(np.abs(np.real(Z[j])))
        return S


def run_plan(plan, f, drivers=None, values=None):
    f = np.asarray(f, dtype=np.float64)
    consts, prefs, tables, samples = pl.resolve_run(plan, f, drivers, values)
    return np.stack([Interp(plan, f, consts, prefs, tables, samples[s]).solve() for s in range(samples.shape[0])])


def dump_plan(plan, f, drivers=None, values=None):
    f = np.asarray(f, dtype=np.float64)
    consts, prefs, tables, samples = pl.resolve_run(plan, f, drivers, values)
    return Interp(plan, f, consts, prefs, tables, samples[0]).matrix()
