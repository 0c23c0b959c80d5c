"""Result containers returned by `run_sp` and the Monte-Carlo / sweep drivers.

Same public surface as the reference's `heavi/sparam.py` (file:line cited per class): a user of
`Sparameters`, `StatSparameters`, `StatSparam`, `NDSparameters` or `frange` sees the same
attributes, shapes and dtypes.  `S[j, i, f]` is the response at port j+1 to an excitation at port
i+1, C-contiguous complex128 with frequency the fastest axis; `f` is a float32 copy of the input
frequencies (network.py:407,421 -- a quirk that is kept).
"""
from __future__ import annotations

import re

import numpy as np
from scipy.interpolate import RegularGridInterpolator


def frange(fmin: float, fmax: float, n: int) -> np.ndarray:
    """Linearly spaced frequency axis (sparam.py:30-41)."""
    return np.linspace(fmin, fmax, n)


_SIJ = re.compile(r"^S(\d+)(\d+)$")


class Sparameters:
    """S-matrix over frequency (sparam.py:49-195)."""

    def __init__(self, S: np.ndarray, f: np.ndarray):
        self._S = S
        self.f = f
        self.nports = S.shape[0]
        self.nfreqs = S.shape[2]

    @property
    def shape(self) -> tuple:
        return self._S.shape

    def S(self, p1: int, p2: int):
        """S[p1, p2], ports counted from 1."""
        for p in (p1, p2):
            if p < 1 or p > self.nports:
                raise ValueError(f"Port {p} out of range")
        return self._S[p1 - 1, p2 - 1, :]

    def __call__(self, p1: int, p2: int):
        return self.S(p1, p2)

    def __getattr__(self, name):
        # greedy first group on purpose: "S110" -> (11, 0), like the reference (sparam.py:83)
        m = _SIJ.match(name)
        if m:
            return self.S(int(m.group(1)), int(m.group(2)))
        raise AttributeError(f"'{self.__class__.__name__}' object has no attribute '{name}'")

    def __dir__(self):
        return [f"S{m}{n}" for m in range(1, 6) for n in range(1, 5)] + list(super().__dir__())


class Interpolator:
    """Callable wrapper around scipy's RegularGridInterpolator (sparam.py:198-213)."""

    def __init__(self, interp: RegularGridInterpolator):
        self._interp = interp

    def __call__(self, *args) -> np.ndarray:
        return np.squeeze(self._interp(np.meshgrid(*args, indexing="ij")))


class StatSparam:
    """One S_ij over Monte-Carlo trials, shape [ntrials, nF] (sparam.py:215-243)."""

    def __init__(self, S: np.ndarray):
        self._S = S
        self.ntrials = S.shape[0]

    @property
    def amplitude_mean(self):
        return np.abs(self._S).mean(axis=0)

    @property
    def amplitude_std(self):
        return np.abs(self._S).std(axis=0)

    @property
    def power_mean(self):
        return (np.abs(self._S) ** 2.0).mean(axis=0)

    @property
    def power_std(self):
        return (np.abs(self._S) ** 2.0).std(axis=0)

    @property
    def rms_mean(self):
        return np.sqrt(np.abs(self._S) ** 2.0).mean(axis=0)

    @property
    def rms_std(self):
        return np.sqrt(np.abs(self._S) ** 2.0).std(axis=0)


class StatSparameters(Sparameters):
    """Per-trial S data of a Monte-Carlo run (sparam.py:246-295).

    `add` takes one trial at a time like the reference; `add_batch` takes the [nS,P,P,nF] block a
    batched GPU run produces without copying it trial by trial."""

    def __init__(self, Ntrials: int):
        self.Ntrials = Ntrials
        self._Sdata: list[np.ndarray] = []

    @property
    def nports(self) -> int:
        return self._Sdata[0].shape[0] if self._Sdata else 0

    @property
    def nfreqs(self) -> int:
        return self._Sdata[0].shape[2] if self._Sdata else 0

    def add(self, S):
        if isinstance(S, Sparameters):
            S = S._S
        if self._Sdata and S.shape != (self.nports, self.nports, self.nfreqs):
            raise ValueError(
                f"Invalid S-parameter shape. Expected {(self.nports, self.nports, self.nfreqs)}, got {S.shape}")
        self._Sdata.append(S)

    def add_batch(self, S4: np.ndarray):
        for k in range(S4.shape[0]):
            self.add(S4[k])

    def S(self, p1: int, p2: int) -> StatSparam:
        return StatSparam(np.stack([S[p1 - 1, p2 - 1, :] for S in self._Sdata], axis=0))

    def __call__(self, p1: int, p2: int) -> StatSparam:
        return self.S(p1, p2)

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return Sparameters.__getattr__(self, name)


class StatMoments:
    """The six StatSparam statistics of one S_ij, already reduced over the trials (on the GPU)."""

    def __init__(self, moments: np.ndarray, ntrials: int):
        self._m = moments            # [6, nF]
        self.ntrials = ntrials

    amplitude_mean = property(lambda self: self._m[0])
    amplitude_std = property(lambda self: self._m[1])
    power_mean = property(lambda self: self._m[2])
    power_std = property(lambda self: self._m[3])
    rms_mean = property(lambda self: self._m[4])
    rms_std = property(lambda self: self._m[5])


class DeviceStatSparameters:
    """`StatSparameters` look-alike for a batched Monte-Carlo run whose per-trial S never left the
    GPU: `S(p1, p2)` returns the same six statistics `StatSparam` offers (sparam.py:215-243)."""

    def __init__(self, moments: np.ndarray, ntrials: int, f: np.ndarray):
        self._moments = moments      # [6, P, P, nF]
        self.Ntrials = ntrials
        self.f = f
        self.nports = moments.shape[1]
        self.nfreqs = moments.shape[3]

    def S(self, p1: int, p2: int) -> StatMoments:
        for p in (p1, p2):
            if p < 1 or p > self.nports:
                raise ValueError(f"Port {p} out of range")
        return StatMoments(self._moments[:, p1 - 1, p2 - 1, :], self.Ntrials)

    def __call__(self, p1: int, p2: int) -> StatMoments:
        return self.S(p1, p2)


class NDSparameters(Sparameters):
    """S data on an N-dimensional parameter grid, shape (..., P, P, nF) (sparam.py:299-358)."""

    def __init__(self, mdim_data: np.ndarray, axes: list):
        self._sdata = mdim_data
        self._S = mdim_data
        self._axes = axes
        self._shape = mdim_data.shape
        self._extra_dims = len(self._shape) - 3
        self.nports = self._shape[-2]
        self.nfreqs = self._shape[-1]
        self.f = axes[-1] if axes else None

    @property
    def _outer_axes(self) -> tuple:
        return tuple(self._axes[:self._extra_dims])

    def S(self, p1: int, p2: int) -> np.ndarray:
        """S[..., p1, p2, :] over the whole parameter grid, ports counted from 1."""
        for p in (p1, p2):
            if p < 1 or p > self.nports:
                raise ValueError(f"Port {p} out of range")
        return np.squeeze(self._sdata[(slice(None),) * self._extra_dims + (p1 - 1, p2 - 1, None)])

    def __call__(self, i1: int, i2: int) -> np.ndarray:
        return self.S(i1, i2)

    def intS(self, p1: int, p2: int) -> Interpolator:
        """Linear interpolator over (first parameter of every swept dimension..., f)."""
        axes = [np.squeeze(np.asarray(ax)[0, :]) for ax in self._outer_axes] + [self._axes[-1]]
        return Interpolator(RegularGridInterpolator(axes, self.S(p1, p2), bounds_error=False, fill_value=None))
