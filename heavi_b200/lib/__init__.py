"""Part libraries (host-side netlist builders): `smd`, `tl`, `libgen`, `touchstone`."""
from . import libgen, smd, tl, touchstone  # noqa: F401
