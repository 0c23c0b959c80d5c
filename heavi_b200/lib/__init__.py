"""Part libraries (host-side netlist builders) and the optimiser: `smd`, `tl`, `libgen`, `touchstone`, `emerge`, `mc`, `optim`."""
from . import emerge, libgen, mc, optim, smd, subcircuit, tl, touchstone  # noqa: F401
