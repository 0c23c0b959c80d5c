"""Part libraries (host-side netlist builders) and the optimiser: `smd`, `tl`, `libgen`, `touchstone`, `optim`."""
from . import libgen, optim, smd, subcircuit, tl, touchstone  # noqa: F401
