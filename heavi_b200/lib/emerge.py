"""N-port block from in-memory S data (an EMerge field-solver export): the array counterpart of a
Touchstone file.

Behavioural equivalent of the reference's `heavi/lib/emerge.py:14-60`: `EMergeModel((F, S))` with
`F[nf]` in Hz and `S[nf, P, P]`, 50 ohm reference, one `interp1d(kind=interp_type,
bounds_error=False, fill_value=(S[0], S[-1]))` per S_ij, stamped through `Network.n_port_S`.  The
traces are `SplineTrace`s, so the device evaluates the splines itself.  (`renormalize` works here; the
reference's version reads an attribute it never sets.)
"""
from __future__ import annotations

import numpy as np

from .libgen import SubCircuit
from .touchstone import SplineTrace


class EMergeModel(SubCircuit):
    def __init__(self, emerge_data: tuple[np.ndarray, np.ndarray], interp_type: str = "cubic"):
        super().__init__()
        self.Fdata, self.Sdata = emerge_data
        self.Fdata = np.asarray(self.Fdata, dtype=np.float64)
        self.Sdata = np.array(self.Sdata, dtype=np.complex128)
        if self.Sdata.ndim != 3 or self.Sdata.shape[1] != self.Sdata.shape[2] or self.Sdata.shape[0] != self.Fdata.shape[0]:
            raise ValueError(f"S data must have shape (n_frequencies, P, P); got {self.Sdata.shape} for {self.Fdata.shape[0]} frequencies")
        self.n_ports: int = self.Sdata.shape[1]
        self.n_nodes = self.n_ports
        self.n_frequencies = self.Sdata.shape[0]
        self.interp_type: str = interp_type
        self.refimp: float = 50

    def renormalize(self, new_impedance: float) -> "EMergeModel":
        """Re-reference the S data to a new real impedance (same transform as touchstone.py:233-251)."""
        z_old, z_new = self.refimp, new_impedance
        P = self.n_ports
        A = np.eye(P, P) * np.sqrt(z_new / z_old) * (1 / (z_new + z_old))
        R = np.eye(P, P) * ((z_new - z_old) / (z_new + z_old))
        iA = np.linalg.pinv(A)
        E = np.eye(P, P)
        for k in range(self.n_frequencies):
            S = self.Sdata[k, :, :]
            self.Sdata[k, :, :] = iA @ (S - R) @ np.linalg.pinv(E - R @ S) @ A
        self.refimp = new_impedance
        return self

    def __on_connect__(self):
        traces = [[SplineTrace(self.Fdata, self.Sdata[:, i, j], self.interp_type) for j in range(self.n_ports)]
                  for i in range(self.n_ports)]
        self.network.n_port_S(self.network.gnd, [self.nodes[i] for i in range(1, self.n_ports + 1)],
                              traces, self.refimp)
