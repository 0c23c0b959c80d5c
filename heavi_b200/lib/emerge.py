"""N-port block from in-memory S data (an EMerge field-solver export): the array counterpart of a
Touchstone file.

Behaviour of the reference's `heavi/lib/emerge.py:14-60`: `EMergeModel((F, S))` with `F[nf]` in Hz and
`S[nf, P, P]`, 50 ohm reference, one cubic `interp1d` with constant end-value fill per S_ij, stamped
through `Network.n_port_S`.  Parsing aside it shares everything with `FileBasedNPort` (spline traces the
device evaluates itself, renormalisation), so it is little more than a constructor here.
(`renormalize` works; the reference's version reads an attribute it never sets.)
"""
from __future__ import annotations

import numpy as np

from .libgen import SubCircuit
from .touchstone import renormalize_s, spline_traces


class EMergeModel(SubCircuit):
    def __init__(self, emerge_data: tuple[np.ndarray, np.ndarray], interp_type: str = "cubic"):
        super().__init__()
        freqs, sdata = emerge_data
        self.Fdata = np.asarray(freqs, dtype=np.float64)
        self.Sdata = np.array(sdata, dtype=np.complex128)
        ok = self.Sdata.ndim == 3 and self.Sdata.shape[1] == self.Sdata.shape[2] and self.Sdata.shape[0] == self.Fdata.shape[0]
        if not ok:
            raise ValueError(f"S data must have shape (n_frequencies, P, P); got {self.Sdata.shape} "
                             f"for {self.Fdata.shape[0]} frequencies")
        self.n_frequencies, self.n_ports = self.Sdata.shape[0], self.Sdata.shape[1]
        self.n_nodes = self.n_ports
        self.interp_type = interp_type
        self.refimp: float = 50

    def renormalize(self, new_impedance: float) -> "EMergeModel":
        self.Sdata = renormalize_s(self.Sdata, self.refimp, new_impedance)
        self.refimp = new_impedance
        return self

    def __on_connect__(self):
        terminals = [self.nodes[i] for i in range(1, self.n_ports + 1)]
        self.network.n_port_S(self.network.gnd, terminals, spline_traces(self.Fdata, self.Sdata, self.interp_type),
                              self.refimp)
