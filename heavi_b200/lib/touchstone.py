"""Touchstone (v1) import: file -> S_ij(f) cubic splines -> `Network.n_port_S`.

Behavioural equivalent of the reference's `heavi/lib/touchstone.py:102-263`: same option-line
defaults (GHz, S, MA, 50 ohm), same "odd count starts a new frequency" record assembly, the
2-port-only transpose (:229-230) and `interp1d(kind='cubic', bounds_error=False,
fill_value=(S[0], S[-1]))` per S_ij (:253-260).  Each spline is wrapped in `SplineTrace`, which is
still a plain callable of f but also exposes the B-spline (knots, coefficients, degree, end
values) so that the plan compiler can evaluate it on the device instead of tabulating 64 traces
on the host.
"""
from __future__ import annotations

import re
from pathlib import Path

import numpy as np
from scipy.interpolate import interp1d

from .libgen import SubCircuit

try:
    from loguru import logger
except Exception:                       # pragma: no cover
    import logging
    logger = logging.getLogger("heavi_b200")

_FUNIT = {"hz": 1, "khz": 1000, "mhz": 1e6, "ghz": 1e9}
_SUFFIXES = (".s1p", ".s2p", ".s3p", ".s4p", ".snp", ".ts")


def _ma_ri(mag, angle):
    return mag * np.exp(1j * np.radians(angle))


def _db_ri(db, angle):
    return 10 ** (db / 20) * np.exp(1j * np.radians(angle))


def _ri_ri(real, imag):
    return real + 1j * imag


_DATAMAP = {"ma": _ma_ri, "db": _db_ri, "ri": _ri_ri}


class SplineTrace:
    """One S_ij(f): scipy `interp1d` with constant end-value fill, plus its B-spline data."""

    def __init__(self, x: np.ndarray, y: np.ndarray, kind: str = "cubic"):
        self.x = np.asarray(x, dtype=np.float64)
        self.y = np.asarray(y, dtype=np.complex128)
        self.kind = kind
        self._interp = interp1d(self.x, self.y, kind=kind, bounds_error=False, fill_value=(self.y[0], self.y[-1]))

    def __call__(self, f):
        return self._interp(f)

    def bspline(self):
        """(t, c, k) of the underlying scipy BSpline, or None if this is not a spline kind."""
        spl = getattr(self._interp, "_spline", None)
        if spl is None or not hasattr(spl, "t"):
            return None
        return (np.asarray(spl.t, dtype=np.float64), np.asarray(spl.c, dtype=np.complex128).reshape(len(spl.c)),
                int(spl.k))


def renormalize_s(sdata: np.ndarray, z_old: float, z_new: float) -> np.ndarray:
    """S[nf, P, P] referenced to `z_old` -> referenced to `z_new` (both real, equal on all ports):
    S' = A^-1 (S - R)(I - R S)^-1 A with A = sqrt(z_new / z_old) / (z_new + z_old) I and
    R = (z_new - z_old)/(z_new + z_old) I, the transform of touchstone.py:233-251."""
    P = sdata.shape[1]
    eye = np.eye(P, P)
    A = eye * np.sqrt(z_new / z_old) * (1 / (z_new + z_old))
    R = eye * ((z_new - z_old) / (z_new + z_old))
    iA = np.linalg.pinv(A)
    out = np.empty_like(sdata)
    for k in range(sdata.shape[0]):
        S = sdata[k, :, :]
        out[k, :, :] = iA @ (S - R) @ np.linalg.pinv(eye - R @ S) @ A
    return out


def spline_traces(fdata: np.ndarray, sdata: np.ndarray, kind: str):
    """P x P nested list of `SplineTrace`s, one per S_ij(f)."""
    P = sdata.shape[1]
    return [[SplineTrace(fdata, sdata[:, i, j], kind) for j in range(P)] for i in range(P)]


class FileBasedNPort(SubCircuit):
    """N-port defined by a Touchstone file.  `connect(n1, ..., nP)` stamps it."""

    def __init__(self, filename: str, n_ports: int = None, ignore_extension: bool = False,
                 interp_type: str = "cubic"):
        super().__init__()
        path = Path(filename)
        if not path.exists():
            raise ValueError(f"The provided filename {path} does not exist.")
        if path.suffix not in _SUFFIXES and not ignore_extension:
            raise ValueError(
                f"Can only open touchstone files, files with {path.suffix} are not a valid touchstone extension.")
        with open(filename, "r") as fh:
            self._lines = fh.read().split("\n")
        self.freq_unit: float = 1e9
        self.param_type: str = "s"
        self.data_type: str = "ma"
        self._converter = _DATAMAP["ma"]
        self.refimp: float = 50
        self.n_ports: int = None
        self.interp_type: str = interp_type
        self.Sdata: np.ndarray = None
        self.Fdata: np.ndarray = None
        self._parse_touchstone()

    def summarize_data(self):
        print(f"Frequency unit: {self.freq_unit}")
        print(f"Parameter type: {self.param_type}")
        print(f"Data type: {self.data_type}")
        print(f"Converter: {self._converter}")
        print(f"Reference impedance: {self.refimp}")
        print(f"Number of ports: {self.n_ports}")

    def _parse_touchstone(self) -> None:
        records, freqs = [], []
        current, freq = [], None
        for raw in self._lines:
            if not raw or not raw.strip():
                continue
            low = raw.strip().lower()
            if low[0] == "!":
                continue
            if low[0] == "#":
                unit = re.findall(r"(hz|khz|mhz|ghz)", low)
                ptype = re.findall(r"([syzgh])", low)
                dtype = re.findall(r"(ma|db|ri)", low)
                ref = re.findall(r"r\s+(\d+)", low)
                if unit:
                    self.freq_unit = _FUNIT[unit[0]]
                if ptype:
                    self.param_type = ptype[0]
                if dtype:
                    self.data_type = dtype[0]
                    self._converter = _DATAMAP[dtype[0]]
                if ref:
                    self.refimp = float(ref[0])
                continue
            if low[0] == "[":
                logger.warning("Version 2.0 touchstone files are not supported yet and keywords will be ignored")
                continue
            nums = [float(tok) for tok in raw.split("!")[0].split()]
            if not nums:
                continue
            if len(nums) % 2 == 1:          # odd count: a frequency followed by data -> new record
                if current:
                    records.append(current)
                    freqs.append(freq)
                freq, current = nums[0], nums[1:]
            else:                           # continuation line of the same record
                current.extend(nums)
        records.append(current)
        freqs.append(freq)

        self.n_ports = int(np.sqrt(len(records[0]) // 2))
        self.n_nodes = self.n_ports
        self.n_frequencies = len(freqs)
        self.Fdata = np.array(freqs) * self.freq_unit
        self.Sdata = np.zeros((self.n_frequencies, self.n_ports, self.n_ports), dtype=complex)
        for i, rec in enumerate(records):
            vals = [self._converter(a, b) for a, b in zip(rec[::2], rec[1::2])]
            mat = np.array(vals).reshape(self.n_ports, self.n_ports)
            if self.n_ports == 2:           # 2-port files list S11 S21 S12 S22
                mat = mat.T
            self.Sdata[i, :, :] = mat

    def renormalize(self, new_impedance: float) -> "FileBasedNPort":
        """Re-reference the S data to a new real impedance (touchstone.py:233-251)."""
        self.Sdata = renormalize_s(self.Sdata, self.refimp, new_impedance)
        self.refimp = new_impedance
        return self

    def __on_connect__(self):
        terminals = [self.nodes[i] for i in range(1, self.n_ports + 1)]
        self.network.n_port_S(self.network.gnd, terminals, spline_traces(self.Fdata, self.Sdata, self.interp_type),
                              self.refimp)
