"""Quarter-wave transmission-line parts: power divider, hybrid couplers, multi-section transformer.

Behaviour follows the reference's `heavi/lib/tl.py:7-112`; the topology of every part is written down
as a table of arms and expanded by one shared routine into `Network.TL` / `Network.resistor` calls, in
the reference's order (the order fixes node and component numbering, and with it the stamp table).
"""
from __future__ import annotations

import math

import numpy as np

from ..filters import impedance_transformer_cheb
from ..params import LinePhase, PhaseConstant, enforce_simparam
from .libgen import FourNodeSubCircuit, ThreeNodeSubCircuit, TwoNodeSubCircuit

_C0 = 299792458
_SQRT2 = math.sqrt(2.0)


def _quarter_wave(fc: float, er: float) -> float:
    """Physical length of a quarter wavelength at `fc` in a medium of relative permittivity `er`."""
    return _C0 / (fc * np.sqrt(er)) / 4


class _ArmTable:
    """Mixin: a part described by `ARMS`, rows of (terminal a, terminal b, quarter waves, m, d) for a line
    of impedance Z0 * m / d, and `LOADS`, rows of (terminal a, terminal b, R / Z0).  Terminals are the
    part's own node numbers.  (Multiplier and divisor are kept apart so that Z0 / sqrt(2) is rounded the
    way the reference rounds it.)"""

    ARMS: tuple = ()
    LOADS: tuple = ()

    def _configure(self, Z0, fc, er):
        self.Z0, self.fc, self.er = Z0, fc, er

    def __on_connect__(self):
        unit = _quarter_wave(self.fc, self.er)
        beta = LinePhase(self.er)                 # one object for all arms: evaluated once, closed form on the device
        net = self.network
        for a, b, quarters, mul, div in self.ARMS:
            net.TL(self.node(a), self.node(b), beta, quarters * unit, self.Z0 * mul / div)
        for a, b, rrel in self.LOADS:
            net.resistor(self.node(a), self.node(b), self.Z0 * rrel)


class Wilkinson(_ArmTable, ThreeNodeSubCircuit):
    """In-phase two-way divider: two sqrt(2) Z0 quarter-wave arms and a 2 sqrt(2) Z0 isolation load
    (the reference's values, lib/tl.py:17-27)."""

    ARMS = ((1, 2, 1, _SQRT2, 1.0), (1, 3, 1, _SQRT2, 1.0))
    LOADS = ((2, 3, 2 * _SQRT2),)

    def __init__(self, Z0, fc, er=1):
        super().__init__()
        self._configure(Z0, fc, er)


class BranchlineCoupler(_ArmTable, FourNodeSubCircuit):
    """90-degree hybrid.  The fourth arm runs from terminal 3 to terminal 3, as in the reference
    (lib/tl.py:45): kept on purpose, it is what exercises the duplicate-index stamp rule
    (SURVEY.md App. A)."""

    ARMS = ((1, 2, 1, 1.0, _SQRT2), (4, 3, 1, 1.0, _SQRT2), (1, 4, 1, 1.0, 1.0), (3, 3, 1, 1.0, 1.0))

    def __init__(self, Z0, fc, er=1):
        super().__init__()
        self._configure(Z0, fc, er)


class RatraceCoupler(_ArmTable, FourNodeSubCircuit):
    """180-degree hybrid ring: three quarter-wave arms and one of three quarter waves, all sqrt(2) Z0."""

    ARMS = ((1, 2, 1, _SQRT2, 1.0), (2, 3, 1, _SQRT2, 1.0), (3, 4, 1, _SQRT2, 1.0), (1, 4, 3, _SQRT2, 1.0))

    def __init__(self, Z0, fc, er=1):
        super().__init__()
        self._configure(Z0, fc, er)

    # named terminals
    sum = property(lambda self: self.node(2))
    diff = property(lambda self: self.node(4))
    delta = diff
    p1 = property(lambda self: self.node(1))
    p2 = property(lambda self: self.node(3))


class Transformer(TwoNodeSubCircuit):
    """Chebyshev multi-section quarter-wave transformer between Z1 and Z2."""

    def __init__(self, Z1, Z2, fc, N_sections: int = 1, ripple=0.05, er=1):
        super().__init__()
        self.Z1, self.Z2, self.fc = Z1, Z2, fc
        self.N_sections, self.ripple, self.er = N_sections, ripple, er

    def __on_connect__(self):
        sections = impedance_transformer_cheb(self.Z1, self.Z2, self.ripple, self.N_sections)[1:-1]
        length = _quarter_wave(self.fc, self.er)
        inner = [self.network.node() for _ in range(self.N_sections - 1)]
        stops = [self.node(1), *inner, self.node(2)]
        beta = PhaseConstant(enforce_simparam(self.er))          # 2 pi f / c0 * sqrt(er), the reference's order here
        for left, right, z0 in zip(stops, stops[1:], sections):
            self.network.TL(left, right, beta, length, float(z0))
