"""Randomised two-port for Monte-Carlo S-parameter runs (behaviour of the reference's `heavi/lib/mc.py:7-65`).

A cable-like block whose match, loss and electrical length are drawn once per trial:

    S11 = S22 = exp(j phi) * (v - 1)/(v + 1) * (Zp - Z0)/(Zp + Z0)
    S21 = S12 = sqrt(1 - S11^2) * 10^(-att * len / 20) * 10^(dA / 20) * exp(j (k(f) len + dphi pi/180))

with v ~ U(1, VSWR), phi ~ U(0, 2 pi), Zp ~ N(Z0, Z0_error / nsigma), dA ~ N(0, amplitude_stability /
nsigma) [dB], dphi ~ N(0, phase_stability / nsigma) [deg] and k(f) = 2 pi f sqrt(er) / c0.  The five
`Random`s are registered with the `MonteCarlo` in exactly that order, which is what makes a seeded run
reproduce the reference's draws.  S11 and S21 read those Randoms when they are called, so their tables
change from trial to trial; the batched drivers detect that
(`CompiledNetwork._tables_follow_drivers`) and run one sweep per trial, like `for i in mc.iterate(N)`.
"""
from __future__ import annotations

import numpy as np

from ..params import MonteCarlo
from ..transformations import VSWR_to_S11, Z_to_S11
from .libgen import TwoNodeSubCircuit

_C0 = 299792458


class RandomTwoPort(TwoNodeSubCircuit):
    def __init__(self, mcsim: MonteCarlo, length: float, Z0: float, VSWR: float = 1, er: float = 1,
                 attenuation: float = 0, phase_stability: float = 0, amplitude_stability: float = 0,
                 Z0_error: float = 0, nsigma: float = 3):
        super().__init__()
        self.mc: MonteCarlo = mcsim
        self.Z0 = Z0
        self._length, self._er = length, er
        self._loss_db = attenuation * length                     # port-to-port loss of the nominal part
        # registration order == draw order (uniform, uniform, gaussian, gaussian, gaussian)
        self._vswr = mcsim.uniform(1, VSWR)
        self._refl_phase = mcsim.uniform(0, 2 * np.pi)
        self._port_z = mcsim.gaussian(Z0, Z0_error / nsigma)
        self._gain_db = mcsim.gaussian(0, amplitude_stability / nsigma)
        self._phase_deg = mcsim.gaussian(0, phase_stability / nsigma)
        self.fS11 = self.fS22 = self._reflection
        self.fS21 = self.fS12 = self._transmission

    def _reflection(self, f):
        rotate = np.exp(1j * self._refl_phase(f))
        return rotate * VSWR_to_S11(self._vswr(f)) * Z_to_S11(self._port_z(f), self.Z0)

    def _transmission(self, f):
        kz = 2 * np.pi * f / (_C0) * np.sqrt(self._er)
        through = (1 - self._reflection(f) ** 2) ** (0.5)
        fixed, drawn = 10 ** (-(self._loss_db) / 20), 10 ** (self._gain_db(f) / 20)
        # left to right, the reference's order of the four factors
        return through * fixed * drawn * np.exp(1j * (kz * self._length + self._phase_deg(f) * np.pi / 180))

    def __on_connect__(self):
        a, b = self.node(1), self.node(2)
        self.network.n_port_S(self.gnd, [a, b], [[self.fS11, self.fS12], [self.fS21, self.fS22]], self.Z0)
