"""Frequency-callable parameter values and the Monte-Carlo / sweep drivers.

Mirrors the value semantics of the reference's `heavi/numeric.py` (file:line cited per class) so
that user scripts keep working, but every parameter also *describes itself structurally*
(`describe()`), which is what lets `plan.py` freeze a netlist into a stamp table for the GPU
instead of keeping Python closures on the hot path:

    ("const", complex)          value fixed for the whole run          Scalar
    ("sample", owner, op)       one real number per Monte-Carlo sample / sweep point; op in
                                {"id", "neg", "inv"}                    Random, Param (+Negative/Inverse)
    ("table", callable)         arbitrary Python callable of f; tabulated on the host per run
    ("beta", er, c0)            closed form (2 pi f / c0) sqrt(er(f)) evaluated on the device
"""
from __future__ import annotations

from itertools import product
from typing import Callable, Generator

import numpy as np

from .results import NDSparameters, StatSparameters

_C0 = 299792458
_TWOPI = np.float64(2 * np.pi)


class Uninitialized:
    """Placeholder for a Random/Param that has not been drawn yet (numeric.py:82-101)."""

    def __repr__(self):
        return "Uninitialized"

    def _fail(self, *a, **k):
        raise ValueError("Uninitialized value")

    __call__ = __mul__ = __rmul__ = __add__ = __radd__ = __sub__ = __rsub__ = _fail


class SimParam:
    """Base class: `p(f)` -> value at the frequencies f (numeric.py:103-140)."""

    _eval_f = 1e9

    def __init__(self, unit: str = ""):
        self.unit = ""
        self._inf = np.inf

    @property
    def value(self):
        return self(0)

    def scalar(self):
        return self.value

    def initialize(self) -> None:
        pass

    def __call__(self, f):
        return np.nan_to_num(np.ones_like(f) * self._value, posinf=self._inf, neginf=-self._inf)

    def __repr__(self) -> str:
        return f"SimParam({self.value})"

    def negative(self) -> "SimParam":
        return Function(lambda f: -self(f))

    def inverse(self) -> "SimParam":
        return Function(lambda f: 1 / self(f))

    # -- structural description used by plan.py -------------------------------------------
    def describe(self):
        # Unknown subclasses (e.g. an optimiser variable) that keep a mutable scalar in
        # `_value` behave like a one-number-per-run sample; anything else is tabulated.
        if type(self).__call__ is SimParam.__call__ and hasattr(self, "_value"):
            return ("sample", self, "id")
        return ("table", self)


class Scalar(SimParam):
    """Fixed number (numeric.py:142-159)."""

    def __init__(self, value, unit: str = "", inf: float = np.inf):
        self._value = value
        self._inf = inf
        self.unit = unit

    def __repr__(self) -> str:
        return f"SimValue({self._value})"

    def negative(self):
        return Scalar(-self._value)

    def inverse(self):
        return Scalar(1 / self._value)

    def describe(self):
        return ("const", complex(self(np.float64(1.0))))


class _Wrapped(SimParam):
    _op = "id"

    def __init__(self, simparam: SimParam, inf: float = np.inf):
        self._simparam = simparam
        self.unit = simparam.unit
        self._inf = inf

    def describe(self):
        inner = self._simparam.describe()
        if inner[0] == "sample" and inner[2] == "id" and np.isinf(self._inf):
            return ("sample", inner[1], self._op)
        return ("table", self)


class Negative(_Wrapped):
    """-p(f) (numeric.py:162-178)."""
    _op = "neg"

    def __repr__(self):
        return f"Negative({self._simparam})"

    def __call__(self, f):
        return np.nan_to_num(-self._simparam(f), posinf=self._inf, neginf=-self._inf)

    def negative(self):
        return self._simparam


class Inverse(_Wrapped):
    """1/p(f) (numeric.py:180-195)."""
    _op = "inv"

    def __repr__(self):
        return f"Inverse({self._simparam})"

    def __call__(self, f):
        return np.nan_to_num(1 / self._simparam(f), posinf=self._inf, neginf=-self._inf)

    def inverse(self):
        return self._simparam


class Function(SimParam):
    """Arbitrary callable of the frequency array (numeric.py:197-218)."""

    def __init__(self, function: Callable, unit: str = "", inf: float = np.inf):
        self._func = function
        self._inf = inf
        self.unit = unit

    def __call__(self, f):
        return np.nan_to_num(self._func(f), posinf=self._inf, neginf=-self._inf)

    def __repr__(self):
        return f"Function({self._func})"

    def describe(self):
        return ("table", self)


class PhaseConstant(Function):
    """beta(f) = 2 pi f / c0 * sqrt(er(f)) -- what `Network.transmissionline` wraps in a plain
    `Function` in the reference (network.py:627-633).  Same values when called; additionally
    tells the plan compiler that the device can evaluate it in closed form."""

    def __init__(self, er: SimParam):
        self._er = er
        super().__init__(self._beta)

    def _beta(self, f):
        return _TWOPI * f / _C0 * np.sqrt(self._er(f))

    def describe(self):
        inner = self._er.describe()
        if inner[0] == "const" and np.isinf(self._inf):
            return ("beta", inner[1])
        return ("table", self)


class LinePhase(Function):
    """beta(f) = 2 pi f sqrt(er) / c0 in the operation order of the reference's coupler library
    (lib/tl.py: `2*np.pi*f*np.sqrt(er)/(299792458)`, i.e. ((2 pi f) sqrt(er)) / c0 -- not the order of
    `Network.transmissionline`, see PhaseConstant).  Callable like any Function; a real, positive `er`
    is also a closed form for the device."""

    def __init__(self, er: float):
        self.er = er
        super().__init__(lambda f: 2 * np.pi * f * np.sqrt(er) / (_C0))

    def describe(self):
        er = self.er
        if isinstance(er, (int, float, np.integer, np.floating)) and er > 0 and np.isinf(self._inf):
            return ("beta_mul", float(np.sqrt(er)))
        return ("table", self)


class SMDImpedance(Function):
    """Z(f) of a surface-mount part with its package parasitics (lib/smd.py:103-294 in the reference).
    Still a plain callable of f -- `function` is the part's own formula -- but it also tells the plan
    compiler which closed form it is, so the device evaluates it per point instead of the host
    tabulating it per run:
        kind 0 (resistor)   (a + j w b) * Zc / (a + j w b + Zc),  Zc = 1 / (j w c)     a = R, b = L, c = C
        kind 1 (inductor)   1 / (1 / (a + j w b) + 1 / (1 / (j w c)))                  a = ESR, b = L, c = Cpar
        kind 2 (capacitor)  a + j w b + 1 / (j w c)                                    a = ESR, b = ESL, c = C
    with w = 2 pi f."""

    def __init__(self, kind: int, a: float, b: float, c: float, function: Callable):
        super().__init__(function)
        self.kind, self.abc = int(kind), (float(a), float(b), float(c))

    def describe(self):
        if np.isinf(self._inf):
            return ("smd", self.kind, self.abc)
        return ("table", self)


class _Drawn(SimParam):
    """Shared part of Random and Param: a scalar `_value` that a driver mutates between runs."""

    def describe(self):
        if np.isinf(self._inf):
            return ("sample", self, "id")
        return ("table", self)

    def negative(self):
        return Negative(self)

    def inverse(self):
        return Inverse(self)


class Random(_Drawn):
    """Monte-Carlo value (numeric.py:221-254)."""

    def __init__(self, randomizer: Callable, unit: str = "", inf: float = np.inf):
        super().__init__()
        self._randomizer = randomizer
        self._value = Uninitialized()
        self._mean = None
        self._std = None
        self._inf = inf
        self.unit = unit

    @property
    def value(self):
        return self._value

    def initialize(self):
        self._value = self._randomizer()

    def __repr__(self):
        return f"Gaussian({self._mean}, {self._std})"


class Param(_Drawn):
    """Swept value (numeric.py:256-311)."""

    def __init__(self, values: np.ndarray, unit: str = "", inf: float = np.inf):
        super().__init__()
        self._values = values
        self._index = 0
        self._value = Uninitialized()
        self._inf = inf
        self.unit = unit

    @staticmethod
    def lin(start: float, stop: float, Nsteps: int) -> "Param":
        return Param(np.linspace(start, stop, Nsteps))

    @staticmethod
    def range(start: float, stop: float, step: float, *args, **kwargs) -> "Param":
        return Param(np.arange(start, stop, step, *args, **kwargs))

    def __len__(self):
        return len(self._values)

    def __repr__(self):
        return f"Param({self._values[0]}, ..., {self._values[-1]})"

    def __call__(self, f):
        return np.nan_to_num(self._value * np.ones_like(f), posinf=self._inf, neginf=-self._inf)

    def set_index(self, index: int):
        self._index = index

    def initialize(self):
        self._value = self._values[self._index]


class ParameterSweep:
    """N-dimensional parameter sweep driver (numeric.py:313-449)."""

    def __init__(self):
        self.sweep_dimensions: list[tuple[Param, ...]] = []
        self.index_series: list[tuple[int, ...]] = []
        self._index = 0
        self._param_buffer: list = []
        self._S_data: list = []
        self._current_index = None
        self._mdim_data = None
        self._fdata = None

    def lin(self, start: float, stop: float) -> "ParameterSweep":
        self._param_buffer.append((start, None, stop))
        return self

    def step(self, start: float, stepsize: float):
        # the reference returns None here (numeric.py:344-360); kept
        self._param_buffer.append((start, stepsize, None))

    def add(self, N: int):
        params = []
        for start, step, stop in self._param_buffer:
            if step is None:
                params.append(Param.lin(start, stop, N))
            elif stop is None:
                params.append(Param.lin(start, start + step * N, N))
        self.add_dimension(*params)
        self._param_buffer = []
        return params[0] if len(params) == 1 else params

    def add_dimension(self, *params: Param):
        self.sweep_dimensions.append(params)

    def submit(self, S, fs) -> "ParameterSweep":
        self._S_data.append((self._current_index, S))
        self._fdata = fs
        return self

    @property
    def S(self) -> NDSparameters:
        if self._mdim_data is None:
            self._mdim_data = self.generate_mdim()
        return self._mdim_data

    def generate_mdim(self) -> NDSparameters:
        shape = [len(dim[0]) for dim in self.sweep_dimensions]
        proto = self._S_data[0][1]
        ports = np.arange(1, proto.shape[0] + 1)
        S = np.zeros(shape + list(proto.shape), dtype=np.complex128)
        for index, s in self._S_data:
            S[index] = s._S
        axes = [np.array([p._values for p in dim]) for dim in self.sweep_dimensions]
        return NDSparameters(S, axes + [ports, ports, self._fdata])

    def combinations(self) -> list[tuple[int, ...]]:
        total = 1
        for dim in self.sweep_dimensions:
            total *= len(dim)
        if total > 10000:
            raise ValueError(f"Total iterations ({total}) is above 10,000, are you sure you want to continue?")
        lengths = [len(dim[0]) for dim in self.sweep_dimensions]
        return list(product(*[range(n) for n in lengths]))

    def iterate(self) -> Generator:
        self._S_data = []
        for ixs in self.combinations():
            values = []
            self._current_index = ixs
            for i, params in zip(ixs, self.sweep_dimensions):
                for p in params:
                    p.set_index(i)
                    p.initialize()
                    values.append(p._value)
            yield ixs, tuple(values)

    @property
    def params(self) -> list[Param]:
        return [p for dim in self.sweep_dimensions for p in dim]


class MonteCarlo:
    """Monte-Carlo driver (numeric.py:451-507).  Draws come from the legacy global `np.random`
    stream, one scalar per Random per sample, sample-major -- `draw(N)` pre-draws a whole batch
    in exactly that order so that a batched GPU run is seed-compatible with the serial loop."""

    def __init__(self):
        self._random_numbers: list[Random] = []
        self.S: StatSparameters | None = None

    def gaussian(self, mean: float, std: float) -> Random:
        r = Random(lambda: np.random.normal(mean, std))
        r._mean, r._std, r._value = mean, std, mean
        r._kind = "gaussian"
        self._random_numbers.append(r)
        return r

    def uniform(self, low: float, high: float) -> Random:
        r = Random(lambda: np.random.uniform(low, high))
        self._random_numbers.append(r)
        return r

    def iterate(self, N: int) -> Generator[int, None, None]:
        self.S = StatSparameters(N)
        for i in range(N):
            for r in self._random_numbers:
                r.initialize()
            yield i

    def draw(self, N: int) -> np.ndarray:
        """[N, nRandom] matrix of draws in the order `iterate(N)` would make them."""
        rs = self._random_numbers
        if rs and all(getattr(r, "_kind", None) == "gaussian" for r in rs):
            # np.random.normal(m, s) is m + s * legacy_gauss(); a sized standard_normal call walks the
            # same legacy stream in the same order, so this is bit-identical to the scalar loop
            z = np.random.standard_normal(size=(N, len(rs)))
            mean = np.array([r._mean for r in rs], dtype=np.float64)
            std = np.array([r._std for r in rs], dtype=np.float64)
            return mean[None, :] + std[None, :] * z
        out = np.empty((N, len(self._random_numbers)), dtype=np.float64)
        for i in range(N):
            for k, r in enumerate(self._random_numbers):
                out[i, k] = r._randomizer()
        return out

    def submit_S(self, S) -> None:
        self.S.add(S)


def enforce_simparam(value, inverse: bool = False, unit: str = "") -> SimParam:
    """number / callable / SimParam -> SimParam (numeric.py:510-539)."""
    if isinstance(value, SimParam):
        return value.inverse() if inverse else value
    if callable(value):
        if inverse:
            return Function(lambda f: 1 / value(f))
        return Function(value, unit=unit)
    if isinstance(value, (int, float, complex)):
        return Scalar(1 / value, unit=unit) if inverse else Scalar(value, unit=unit)
    raise ValueError(f"Invalid value type: {type(value)}")


def set_print_frequency(frequency: float) -> None:
    """numeric.py:560-575, including its inverted range check (App. B of SURVEY.md)."""
    if not isinstance(frequency, (int, float)):
        raise ValueError("Frequency must be a float")
    if 0 < frequency < 1e15:
        raise ValueError("Frequency must be greater than 0 and less than 1e15 Hz")
    SimParam._eval_f = frequency
