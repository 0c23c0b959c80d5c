"""S-parameter optimisation helpers: variables, band requirements and objective functions.

Same surface as the reference's `heavi/lib/optim.py` (`OptimVar` :31-64, `PNorm` :66-100,
`dBbelow`/`dBabove` :120-172, `FrequencyLevel` :187-236, `Optimiser` :239-438).  The reference's
objective calls `network.run_sp(fs)` once per candidate (:427,:434); `generate_objective` keeps
that contract, and `generate_population_objective` evaluates a whole population -- e.g. one
generation of `scipy.optimize.differential_evolution(..., vectorized=True)` -- as ONE batched GPU
run in which every candidate is a sample.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from enum import Enum
from typing import Callable

import numpy as np

from ..params import SimParam
from ..results import Sparameters


class OptimVar(SimParam):
    """Optimisation variable; `mapping` (e.g. 10**x) translates the optimiser's coordinate into the
    physical value the circuit sees."""

    def __init__(self, initial_value: float, bounds: tuple[float, float] | None = None,
                 mapping: Callable | None = None, unit: str = ""):
        if mapping is not None and not callable(mapping):
            raise TypeError(f"The variable mapping must be a callable function. Use dBbelow or dBabove. Not {type(mapping)}")
        self._coord = initial_value
        self.bounds = bounds
        self.mapping = mapping
        self.unit = unit
        self._inf = np.inf
        self._physical = None     # set by the batched drivers while they probe / replay single candidates

    # `_value` is what the plan compiler reads as this variable's per-run sample value
    @property
    def _value(self):
        if self._physical is not None:
            return self._physical
        return self._coord if self.mapping is None else self.mapping(self._coord)

    @_value.setter
    def _value(self, v):          # batched drivers write the last sample back; the coordinate rules here
        pass

    def _set_physical(self, v):
        """Batched drivers: make the variable read `v` (a physical value, i.e. after the mapping) until
        `_set_physical(None)` -- the hook engine._tables_follow_drivers / the per-candidate fallback use."""
        self._physical = None if v is None else float(v)

    def set(self, value: float):
        self._coord = value

    @property
    def value(self):
        return self(0)

    @value.setter
    def value(self, v):
        self._coord = v

    def __call__(self, f):
        return self._value * np.ones_like(f)

    def describe(self):
        return ("sample", self, "id")


class PNorm(Enum):
    MIN = 0
    MAX = 1
    P1 = 2
    P2 = 3

    def get_metric(self) -> Callable:
        if self is PNorm.MAX:
            return np.max
        if self is PNorm.MIN:
            return np.min
        if self is PNorm.P1:
            return np.mean
        return lambda x: np.sqrt(np.mean(x ** 2))


def generalized_mean(norm: float) -> Callable:
    return lambda x: np.mean(x ** norm) ** (1 / norm)


def _meval(norm):
    return norm.get_metric() if isinstance(norm, PNorm) else generalized_mean(norm)


def dBbelow(dB_level: float, norm: float | PNorm = PNorm.P2) -> Callable:
    """Penalty for |S| (dB) rising above `dB_level`."""
    meval = _meval(norm)
    return lambda S: meval(np.clip(20 * np.log10(np.abs(S)) - dB_level, a_min=0, a_max=None)) / 10


def dBabove(dB_level: float, norm: float | PNorm = PNorm.P2) -> Callable:
    """Penalty for |S| (dB) falling below `dB_level`."""
    meval = _meval(norm)
    return lambda S: meval(np.clip(dB_level - 20 * np.log10(np.abs(S)), a_min=0, a_max=None)) / 10


class Requirement(ABC):
    @abstractmethod
    def eval(self, S: Sparameters) -> float:
        ...


class FrequencyLevel(Requirement):
    """|S_ij| above / below a level over nF points of [fmin, fmax]."""

    def __init__(self, fmin: float, fmax: float, nF: int, sparam: tuple[int, int], upper_limit: Callable | None = None,
                 lower_limit: Callable | None = None, weight: float = 1, f_norm: float | PNorm = 4):
        self.fs = np.linspace(fmin, fmax, nF)
        self.param = sparam
        self.slc: slice = None
        self.weight = weight
        self.upper_limit, self.lower_limit = upper_limit, lower_limit
        if upper_limit is not None and lower_limit is not None:
            self.metric = lambda S: upper_limit(S) + lower_limit(S)
        elif upper_limit is not None:
            self.metric = upper_limit
        elif lower_limit is not None:
            self.metric = lower_limit
        else:
            raise ValueError("At least one of upper_limit or lower_limit must be specified")

    def eval(self, S: Sparameters) -> float:
        return self.metric(S.S(self.param[0], self.param[1])[self.slc]) * self.weight

    def generate_fill_area(self):
        lower = 0 if self.lower_limit is None else self.lower_limit
        upper = -100 if self.upper_limit is None else self.upper_limit
        return (self.fs[0], self.fs[-1], upper, lower)


class Optimiser:
    def __init__(self, network):
        self.network = network
        self.parameters: list[OptimVar] = []
        self.requirements: list[FrequencyLevel] = []

    @property
    def spec_area(self):
        return [req.generate_fill_area() for req in self.requirements]

    @property
    def bounds(self):
        return [p.bounds for p in self.parameters]

    @property
    def x0(self) -> np.ndarray:
        return np.array([p._coord for p in self.parameters])

    def add_param(self, initial, bounds=None, mapping: Callable = None, unit: str = "") -> OptimVar:
        p = OptimVar(initial, bounds=bounds, mapping=mapping, unit=unit)
        self.parameters.append(p)
        return p

    def cap(self, logscale: bool = True, initial: float = -12, limits=(-13, -6)) -> OptimVar:
        if not logscale:
            return self.add_param(10 ** initial, (1 * 10 ** limits[0], 1 * 10 ** limits[1]), unit="f")
        return self.add_param(initial, limits, mapping=lambda x: 10 ** (x), unit="f")

    def ind(self, logscale: bool = True, initial: float = -9, limits=(-12, -5)) -> OptimVar:
        if not logscale:
            return self.add_param(10 ** initial, (10 ** (limits[0]), 10 ** (limits[1])), unit="H")
        return self.add_param(initial, limits, mapping=lambda x: 10 ** (x), unit="H")

    def add_requirement(self, req: Requirement):
        self.requirements.append(req)

    def add_splimit(self, fmin, fmax, nF, sparam, upper_limit=None, lower_limit=None, weight=1, f_norm=4):
        self.add_requirement(FrequencyLevel(fmin, fmax, nF, sparam, upper_limit=upper_limit, lower_limit=lower_limit,
                                            weight=weight, f_norm=f_norm))

    # ------------------------------------------------------------------ objectives
    def _frequency_plan(self, initial):
        if initial is not None:
            for p, v in zip(self.parameters, initial):
                p.set(v)
        fs, n = [], 0
        for req in self.requirements:
            fs = fs + list(req.fs)
            req.slc = slice(n, len(fs))
            n = len(fs)
        return np.array(fs)

    def _cost(self, S: Sparameters, pnorm, dw, dwe) -> float:
        if dw > 0:
            Ms = np.array([abs(req.eval(S)) for req in self.requirements])
            return (np.mean(Ms ** pnorm)) ** (1 / pnorm) + dw * (max(Ms) - min(Ms)) ** dwe
        sub = [abs(req.eval(S)) ** pnorm for req in self.requirements]
        return (sum(sub) / len(self.requirements)) ** (1 / pnorm)

    def generate_objective(self, pnorm=2, initial: np.ndarray = None, differential_weighting: float = 0,
                           differential_weighting_exponent: float = 5) -> Callable:
        """objective(coeffs) -> cost, one `run_sp` per call (optim.py:386-438)."""
        fs = self._frequency_plan(initial)

        def objective(coeffs):
            for p, c in zip(self.parameters, coeffs):
                p.set(c)
            S = self.network.run_sp(fs)
            return self._cost(S, pnorm, differential_weighting, differential_weighting_exponent)

        return objective

    def generate_population_objective(self, pnorm=2, initial: np.ndarray = None, differential_weighting: float = 0,
                                      differential_weighting_exponent: float = 5) -> Callable:
        """objective(X) -> costs for a whole population in one batched GPU run.

        `X` has one candidate per COLUMN (shape [n_vars, n_candidates]), the layout
        `scipy.optimize.differential_evolution(vectorized=True)` passes; a 1-D vector is one candidate."""
        fs = self._frequency_plan(initial)
        f32 = np.array(fs).astype(np.float32)

        def objective(X):
            X = np.asarray(X, dtype=np.float64)
            single = X.ndim == 1
            cand = (X[:, None] if single else X).T                           # [n_candidates, n_vars]
            phys = np.empty_like(cand)
            for k, p in enumerate(self.parameters):
                col = cand[:, k]
                phys[:, k] = col if p.mapping is None else np.array([p.mapping(v) for v in col], dtype=np.float64)
            S4 = self.network.run_sp_batch(fs, self.parameters, phys)
            costs = np.array([self._cost(Sparameters(S4[c], f32), pnorm, differential_weighting,
                                         differential_weighting_exponent) for c in range(S4.shape[0])])
            for p, c in zip(self.parameters, cand[-1]):                      # state after the serial loop
                p.set(c)
            return float(costs[0]) if single else costs

        return objective
