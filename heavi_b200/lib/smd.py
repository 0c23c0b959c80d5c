"""SMD R/L/C parts with package parasitics, each stamped as ONE callable-valued impedance.

Same models and package data as the reference's `heavi/lib/smd.py:53-294`; the impedance
callables use the same operation order so the tabulated Z(f) agrees bit for bit.
"""
from __future__ import annotations

import math
from enum import Enum

from ..params import SMDImpedance
from .libgen import TwoNodeSubCircuit


class SMDResistorSize(Enum):
    R0402 = "0402"
    R0603 = "0603"
    R0805 = "0805"
    R1206 = "1206"


class SMDCapacitorSize(Enum):
    C0402 = "0402"
    C0603 = "0603"
    C0805 = "0805"
    C1206 = "1206"


class SMDInductorSize(Enum):
    L0402 = "0402"
    L0603 = "0603"
    L0805 = "0805"
    L1206 = "1206"


# package -> parasitic pair; the meaning of the pair depends on the part
_R_LC = {"0402": (0.8e-9, 0.2e-12), "0603": (1.0e-9, 0.3e-12), "0805": (1.5e-9, 0.45e-12), "1206": (1.8e-9, 0.6e-12)}
_L_RC = {"0402": (0.15, 0.3e-12), "0603": (0.10, 0.5e-12), "0805": (0.08, 0.8e-12), "1206": (0.06, 1.2e-12)}
_C_RL = {"0402": (0.1, 0.3e-9), "0603": (0.08, 0.5e-9), "0805": (0.07, 0.7e-9), "1206": (0.05, 1.0e-9)}


def _lookup(table, package, default):
    return table.get(getattr(package, "value", None), default)


class _SMDPart(TwoNodeSubCircuit):
    _label, _unit = "", ""

    def _stamp(self, zfun, nominal, **meta):
        self.function = zfun
        comp = self.network.impedance(self.node(1), self.node(2), zfun, display_value=nominal)
        comp.set_metadata(name=f"{self.package}SMD {self._label}", unit=self._unit, value=nominal, **meta)


class SMDResistor(_SMDPart):
    """(R + jwL) parallel to 1/(jwC)."""
    _label, _unit = "Resistor", "Ω"

    def __init__(self, resistance: float, package: SMDResistorSize):
        super().__init__()
        self.n_nodes = 2
        self.resistance = resistance
        self.package = package
        self.inductance, self.capacitance = _lookup(_R_LC, package, (1.0e-9, 0.3e-12))
        self.function = None

    def __on_connect__(self):
        def z_parasitic(f):
            w = 2 * 3.141592653589793 * f
            Zr = self.resistance
            Zl = 1j * w * self.inductance
            Zc = 1 / (1j * w * self.capacitance)
            return (Zr + Zl) * Zc / (Zr + Zl + Zc)
        z = SMDImpedance(0, self.resistance, self.inductance, self.capacitance, z_parasitic)
        self._stamp(z, self.resistance, inductance=self.inductance, capacitance=self.capacitance)


class SMDInductor(_SMDPart):
    """(ESR + jwL) parallel to 1/(jwCpar)."""
    _label, _unit = "Inductor", "H"

    def __init__(self, inductance: float, package: SMDInductorSize):
        super().__init__()
        self.package = package
        self.n_nodes = 2
        self.inductance = inductance
        self.esr, self.parallel_capacitance = _lookup(_L_RC, package, (0.1, 0.5e-12))
        self.function = None

    def __on_connect__(self):
        def z_parasitic(f):
            w = 2 * math.pi * f
            z_series = self.esr + 1j * w * self.inductance
            z_parallel_cap = 1 / (1j * w * self.parallel_capacitance)
            return 1 / (1 / (z_series) + 1 / (z_parallel_cap))
        z = SMDImpedance(1, self.esr, self.inductance, self.parallel_capacitance, z_parasitic)
        self._stamp(z, self.inductance, esr=self.esr, parallel_capacitance=self.parallel_capacitance)


class SMDCapacitor(_SMDPart):
    """ESR + jwESL + 1/(jwC) in series."""
    _label, _unit = "Capacitor", "F"

    def __init__(self, capacitance: float, package: SMDCapacitorSize):
        super().__init__()
        self.package = package
        self.n_nodes = 2
        self.capacitance = capacitance
        self.esr, self.esl = _lookup(_C_RL, package, (0.1, 0.5e-9))
        self.function = None

    def __on_connect__(self):
        def z_parasitic(f):
            w = 2 * math.pi * f
            z_series = self.esr + 1j * w * self.esl
            z_c = 1 / (1j * w * self.capacitance)
            return z_series + z_c
        z = SMDImpedance(2, self.esr, self.esl, self.capacitance, z_parasitic)
        self._stamp(z, self.capacitance, esr=self.esr, esl=self.esl)
