"""Lumped filter synthesis that only ever calls `Network.capacitor/inductor/node`.

Host-side netlist builder (runs once per circuit), kept because BASELINE.json configs 1, 2 and 4
build their circuits with `model.filters.cauer_filter(...)`.  Element values and creation order
follow the reference's `heavi/filtering.py` (file:line cited) so that the node numbering and the
L/C values -- hence the stamp table -- are identical.  The formulas are the standard low-pass
prototype (g-value) ladder transformations.
"""
from __future__ import annotations

from enum import Enum
from math import comb

import numpy as np


class FilterType(Enum):
    CHEBYCHEV = 1
    BUTTERWORTH = 2


class BandType(Enum):
    HIGHPASS = 1
    LOWPASS = 2
    BANDPASS = 3
    BANDSTOP = 4


class CauerType(Enum):
    TYPE1 = 1     # first element in shunt
    TYPE2 = 2     # first element in series


def prototype_chebychev(order: int, ripple: float):
    """g_1..g_N and the load scaler g_{N+1} of the equal-ripple prototype (filtering.py:98-137)."""
    N = order
    B = np.log(1 / np.tanh(ripple * np.log(10) / 40))
    y = np.sinh(B / (2 * N))
    a = [np.sin((2 * k - 1) * np.pi / (2 * N)) for k in range(N + 2)]
    b = [y ** 2 + np.sin(k * np.pi / N) ** 2 for k in range(N + 2)]
    g = [1] * (N + 2)
    for k in range(1, N + 2):
        if k == 1:
            g[k] = 2 * a[1] / y
        elif k <= N:
            g[k] = 4 * a[k - 1] * a[k] / (b[k - 1] * g[k - 1])
        else:
            g[k] = 1 if k % 2 == 0 else 1 / np.tanh(B / 4) ** 2
    return g[1:-1], g[-1]


def prototype_butterworth(order: int):
    """Maximally flat prototype g-values; load scaler is 1 (filtering.py:140-161)."""
    N = order
    return [2 * np.sin(np.pi * (2 * i - 1) / (2 * N)) for i in range(1, N + 1)], 1


def impedance_transformer_cheb(Z1: float, Z2: float, A: float, N: int):
    """Section impedances of an N-section Chebyshev transformer (filtering.py:187-237).

    Uses the closed form of the small-reflection theory: the partial reflection coefficients are
    the cosine-series coefficients of A * T_N(sec(theta_m) cos(theta)).  The reference obtains the
    same coefficients symbolically with sympy."""
    secm = np.cosh(1 / N * np.arccosh((1 / (2 * A)) * np.log(Z2 / Z1)))
    # T_N(s cos t) as a polynomial in x = cos t, then into a cosine series
    cheb = np.polynomial.chebyshev.cheb2poly([0] * N + [1])          # coefficients of x^k in T_N(x)
    series = np.zeros(N + 1)
    for k, ck in enumerate(cheb):
        if ck == 0:
            continue
        # (s cos t)^k = s^k 2^{-k} sum_j C(k,j) cos((k-2j) t)
        for j in range(k + 1):
            series[abs(k - 2 * j)] += ck * secm ** k * comb(k, j) / 2 ** k
    gam = []
    for n in range(N // 2 + 1):
        mult = N - 2 * n
        gam.append(A * series[mult] / 2 if mult != 0 else A * series[0] / 2)
    gammas = gam + list(reversed(gam))[((N + 1) % 2):]
    impedances = [Z1]
    last = Z1
    for gk in gammas[:-1]:
        last = last * np.exp(2.0 * float(gk))
        impedances.append(last)
    impedances.append(Z2)
    return impedances


class Filtering:
    """`model.filters`: adds filter networks to a Network (filtering.py:240-461)."""

    def __init__(self, N):
        self.N = N

    def impedance_transformer(self, gnd, port1, port2, Z01: float, Z02: float, fc: float = None,
                              Nsections: int = 2, ripple: float = 0.05, er: float = 1,
                              frange: tuple[float, float] = None):
        """Multi-section quarter-wave transformer.  (The reference's version passes its arguments
        to `transmissionline` in an obsolete order and fails; this one uses
        the current order `port1, port2, Z0, er, L, gnd`.)"""
        if frange is not None:
            f1, f2 = frange
            fc = 0.5 * (f1 + f2)
            A = np.cosh(Nsections * np.arccosh(1 / np.cos(np.pi / 4 * (2 - (f2 - f1) / fc))))
            ripple = 1 / (2 * A) * np.log(Z02 / Z01)
        l0 = 299792458 / fc * 0.25 * (1 / np.sqrt(er))
        imps = impedance_transformer_cheb(Z01, Z02, ripple, Nsections)[1:-1]
        inner = [self.N.node(index="TransformerNode") for _ in range(len(imps) - 1)]
        chain = [port1] + inner + [port2]
        for Z, n1, n2 in zip(imps, chain[:-1], chain[1:]):
            self.N.transmissionline(n1, n2, Z, er, l0, gnd)
        return imps, ripple

    def cauer_filter(self, gnd, port1, port2, fc: float, bw: float, order: int = 2, ripple: float = 0.05,
                     kind: FilterType = FilterType.BUTTERWORTH, type: BandType = BandType.LOWPASS,
                     Z0: float = 50, cauer_type: CauerType = CauerType.TYPE2,
                     chebychev_correction: bool = False):
        """Cauer (ladder) filter between port1 and port2 (filtering.py:308-461).

        Returns (capacitors, inductors, load impedance scaler)."""
        net = self.N
        pool = "internalFilterNode"
        n_inner = int(np.floor((order - (1 if cauer_type == CauerType.TYPE2 else 2)) / 2)) if cauer_type in (CauerType.TYPE1, CauerType.TYPE2) else 0
        rail = [port1] + [net.node(index=pool) for _ in range(n_inner)] + [port2]

        w_c = 2 * np.pi * fc            # cut-off (LP/HP) and centre (BP/BS) angular frequency
        w_0 = w_c
        d_w = 2 * np.pi * bw
        Q = fc / bw

        if kind == FilterType.BUTTERWORTH:
            g, load = prototype_butterworth(order)
        else:
            g, load = prototype_chebychev(order, ripple)
            if chebychev_correction:
                eps = np.sqrt(10 ** (ripple / 10) - 1)
                k = np.cosh(1 / order * np.arccosh(1 / eps))
                if type is BandType.HIGHPASS or type is BandType.BANDSTOP:
                    d_w, w_c = d_w * k, w_c * k
                else:
                    d_w, w_c = d_w / k, w_c / k
                Q = w_0 / d_w

        # TYPE2: odd elements in series, even elements in shunt; TYPE1 the other way round
        series_first = cauer_type == CauerType.TYPE2
        if not series_first:
            load = 1 / load
        caps, inds = [], []

        for i, gk in enumerate(g):
            k = i + 1
            a = (i + (1 if series_first else 0)) // 2
            na = rail[a]
            nb = rail[a + 1] if a + 1 < len(rail) else None
            is_shunt = (k % 2 == 0) if series_first else (k % 2 == 1)
            if type == BandType.LOWPASS:
                if is_shunt:
                    caps.append(net.capacitor(gnd, na, gk / (w_c * Z0)))
                else:
                    inds.append(net.inductor(na, nb, (gk * Z0) / w_c))
            elif type == BandType.HIGHPASS:
                if is_shunt:
                    inds.append(net.inductor(gnd, na, Z0 / (gk * w_c)))
                else:
                    caps.append(net.capacitor(na, nb, 1 / (w_c * gk * Z0)))
            elif type == BandType.BANDPASS:
                if is_shunt:
                    c_par = Q * gk / (w_0 * Z0)
                    l_par = Z0 / (gk * w_0 * Q)
                    caps.append(net.capacitor(gnd, na, c_par))
                    inds.append(net.inductor(gnd, na, l_par))
                else:
                    mid = net.node(index=pool)
                    c_ser = 1 / (w_0 * gk * Z0 * Q)
                    l_ser = Q * gk * Z0 / w_0
                    caps.append(net.capacitor(na, mid, c_ser))
                    inds.append(net.inductor(mid, nb, l_ser))
            elif type is BandType.BANDSTOP:
                if is_shunt:
                    mid = net.node(index=pool)
                    c_par = gk / (Q * w_0 * Z0)
                    l_par = Q * Z0 / (gk * w_0)
                    caps.append(net.capacitor(gnd, mid, c_par))
                    inds.append(net.inductor(mid, na, l_par))
                else:
                    c_ser = Q / (w_0 * gk * Z0)
                    l_ser = gk * Z0 / (w_0 * Q)
                    caps.append(net.capacitor(na, nb, c_ser))
                    inds.append(net.inductor(na, nb, l_ser))
        return caps, inds, load
