"""Import-path compatibility: the reference keeps the part framework in `heavi.lib.libgen`
(`from heavi.lib.libgen import SubCircuit`); the implementation lives in `subcircuit.py`."""
from .subcircuit import FourNodeSubCircuit, SubCircuit, ThreeNodeSubCircuit, TwoNodeSubCircuit  # noqa: F401
