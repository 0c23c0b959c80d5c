"""heavi_b200 -- B200-native drop-in for heavi's frequency-sweep S-parameter path.

    import heavi_b200 as hv
    m = hv.Model(); ...; S = m.run_sp(hv.frange(1.8e9, 2.2e9, 2001))

The circuit-building API mirrors FennisRobert/heavi; `run_sp` runs on sm_100a CUDA kernels through
the C-ABI library `libheavi_b200.so` (see DESIGN.md / INTEGRATION.md).
"""
from . import lib  # noqa: F401
from .filters import BandType, CauerType, Filtering, FilterType
from .lib.touchstone import FileBasedNPort
from .model import Model
from .netlist import Network, Node, run_sp_many
from .params import Function, MonteCarlo, Param, ParameterSweep, set_print_frequency
from .results import DeviceStatSparameters, NDSparameters, Sparameters, StatSparam, StatSparameters, frange

PI = 3.141592653589793
__all__ = ["Model", "Network", "Node", "frange", "Sparameters", "NDSparameters", "StatSparam", "StatSparameters",
           "FilterType", "BandType", "CauerType", "Filtering", "MonteCarlo", "Param", "Function", "ParameterSweep",
           "set_print_frequency", "FileBasedNPort", "lib", "PI", "run_sp_many"]
