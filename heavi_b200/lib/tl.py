"""Transmission-line couplers built from `Network.TL` (reference: heavi/lib/tl.py:7-112)."""
from __future__ import annotations

import numpy as np

from ..filters import impedance_transformer_cheb
from .libgen import FourNodeSubCircuit, ThreeNodeSubCircuit, TwoNodeSubCircuit

_C0 = 299792458


class _QuarterWavePart:
    def _line_data(self):
        er = self.er
        quarter = _C0 / (self.fc * np.sqrt(er)) / 4

        def beta(f):
            return 2 * np.pi * f * np.sqrt(er) / (_C0)
        return beta, quarter


class Wilkinson(ThreeNodeSubCircuit, _QuarterWavePart):
    def __init__(self, Z0, fc, er=1):
        super().__init__()
        self.Z0, self.fc, self.er = Z0, fc, er

    def __on_connect__(self):
        beta, L = self._line_data()
        Z0 = self.Z0 * np.sqrt(2)
        self.network.TL(self.node(1), self.node(2), beta, L, Z0)
        self.network.TL(self.node(1), self.node(3), beta, L, Z0)
        self.network.resistor(self.node(2), self.node(3), Z0 * 2)


class BranchlineCoupler(FourNodeSubCircuit, _QuarterWavePart):
    def __init__(self, Z0, fc, er=1):
        super().__init__()
        self.Z0, self.fc, self.er = Z0, fc, er

    def __on_connect__(self):
        beta, L = self._line_data()
        Z0 = self.Z0 / np.sqrt(2)
        self.network.TL(self.node(1), self.node(2), beta, L, Z0)
        self.network.TL(self.node(4), self.node(3), beta, L, Z0)
        self.network.TL(self.node(1), self.node(4), beta, L, self.Z0)
        # the reference wires this last arm from node 3 to node 3 (lib/tl.py:45); kept, because
        # it is what exercises the duplicate-index stamp rule
        self.network.TL(self.node(3), self.node(3), beta, L, self.Z0)


class RatraceCoupler(FourNodeSubCircuit, _QuarterWavePart):
    def __init__(self, Z0, fc, er=1):
        super().__init__()
        self.Z0, self.fc, self.er = Z0, fc, er

    def __on_connect__(self):
        beta, L = self._line_data()
        Z0 = self.Z0 * np.sqrt(2)
        self.network.TL(self.node(1), self.node(2), beta, L, Z0)
        self.network.TL(self.node(2), self.node(3), beta, L, Z0)
        self.network.TL(self.node(3), self.node(4), beta, L, Z0)
        self.network.TL(self.node(1), self.node(4), beta, 3 * L, Z0)

    @property
    def sum(self):
        return self.node(2)

    @property
    def diff(self):
        return self.node(4)

    delta = diff

    @property
    def p1(self):
        return self.node(1)

    @property
    def p2(self):
        return self.node(3)


class Transformer(TwoNodeSubCircuit):
    def __init__(self, Z1, Z2, fc, N_sections: int = 1, ripple=0.05, er=1):
        super().__init__()
        self.Z1, self.Z2, self.fc = Z1, Z2, fc
        self.N_sections, self.ripple, self.er = N_sections, ripple, er

    def __on_connect__(self):
        imps = impedance_transformer_cheb(self.Z1, self.Z2, self.ripple, self.N_sections)[1:-1]
        L = _C0 / (self.fc * np.sqrt(self.er)) / 4
        chain = [self.node(1)] + [self.network.node() for _ in range(self.N_sections - 1)] + [self.node(2)]
        for n1, n2, z0 in zip(chain[:-1], chain[1:], imps):
            self.network.TL(n1, n2, lambda f: 2 * np.pi * f / _C0 * np.sqrt(self.er), L, float(z0))
