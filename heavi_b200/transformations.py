"""Reflection-coefficient helpers used by the part libraries (subset of the reference's
`heavi/transformations.py:47-68`; the rest of that module is post-processing outside the S-parameter
path)."""
import numpy as np


def Z_to_S11(Z: np.ndarray, Z0: float = 50) -> np.ndarray:
    """Load impedance -> reflection coefficient."""
    return (Z - Z0) / (Z + Z0)


def VSWR_to_S11(VSWR: np.ndarray) -> np.ndarray:
    """Voltage standing wave ratio -> |reflection coefficient|."""
    return (VSWR - 1) / (VSWR + 1)
