"""Circuit elements as *structural records*.

The reference keeps one class per element whose `SP_matrix_compiler()` returns a Python closure
that adds a dense (k,k,nF) block into the MNA tensor (heavi/base/*.py).  Here an element only
records what the stamp needs -- its kind, its nodes in the reference's per-class `ids` order, and
its parameters -- and `plan.py` freezes the list into the stamp table the CUDA kernels interpret.
The user-visible attributes of the reference classes (`nodes`, `value`, `component_name`, `unit`,
`set_metadata`, `summary`, `Port.index/sig_node/int_node/gnd_node/Z0`) are kept.
"""
from __future__ import annotations

import numpy as np

from .params import SimParam, enforce_simparam

_PREFIX = {-12: "p", -9: "n", -6: "u", -3: "m", 0: "", 3: "k", 6: "M", 9: "T", 12: "P"}


def format_value_units(value, unit: str) -> str:
    """4532, 'Hz' -> '4.53 kHz' (base/component.py:28-47)."""
    exp3 = np.floor(np.log10(np.abs(value)) / 3) * 3
    exp3 = min(12, max(-12, exp3))
    return f"{value / (10 ** exp3):.2f} {_PREFIX[exp3]}{unit}"


class MatSource:
    """Extra MNA row/column requested by an element (node.py:109-117)."""

    def __init__(self, index: int):
        self.index = index

    def __repr__(self):
        return f"MatSource(index={self.index})"


class Element:
    """Base record.  `kind` names the stamp; `sp_sources` extra MNA unknowns for S analysis."""

    kind = "element"
    component_name = "BaseComponent"
    unit = ""
    sp_sources = 0
    PRINT_FREQ = 1e0

    def __init__(self, nodes):
        self.nodes = list(nodes)
        self.sources: list[MatSource] = []
        self.internal_nodes: list = []
        self.nterminals = len(self.nodes)
        self.value = "Undefined"
        self.metadata: dict = {}

    # -- what the reference exposes ----------------------------------------------------------
    @property
    def all_nodes(self):
        return self.nodes + self.internal_nodes

    def set_metadata(self, **kwargs):
        self.metadata.update(kwargs)
        return None            # the reference returns None here as well (component.py:100-101)

    def set_sources(self, sources):
        self.sources = sources

    def indices(self) -> list[int]:
        return [n.index for n in self.nodes] + [s.index for s in self.sources]

    def summary(self) -> str:
        base = dict(name=self.component_name, value=self.value, unit=self.unit)
        base.update(self.metadata)
        head = f'{base.pop("name")} ({format_value_units(base.pop("value"), base.pop("unit"))})'
        lines = [head]
        if base:
            lines.append("-" * len(head))
            lines += [f"    {k}: {v}" for k, v in base.items()]
        return "\n".join(lines)

    def __repr__(self):
        return f"{self.component_name}[{format_value_units(self.value, self.unit)}]"

    # -- what the stamp table needs ----------------------------------------------------------
    def stamp_ids(self) -> list[int]:
        """MNA indices in the reference class's own `mat_slice` order."""
        return [n.index for n in self.nodes]

    def stamp_params(self) -> dict:
        return {}

    def record(self) -> dict:
        """Neutral record understood by the test oracle (oracle/mna_oracle.py)."""
        rec = {"kind": self.kind, "ids": self.stamp_ids()}
        rec.update(self.stamp_params())
        return rec


class Port(Element):
    """Series Z0 between sig and int plus an ideal source int->gnd on an extra row
    (base/port.py:6-42); ids = [sig, int, gnd, src]."""
    kind = "port"
    component_name = "Port"
    unit = "Ohm"
    sp_sources = 1

    def __init__(self, index: int, sig_node, int_node, gnd_node, Z0):
        super().__init__([sig_node, int_node, gnd_node])
        self.index = index
        self.sig_node, self.int_node, self.gnd_node = sig_node, int_node, gnd_node
        self.nterminals = 2
        self.Z0: SimParam = enforce_simparam(Z0)
        self.value = self.Z0.value

    def stamp_ids(self):
        return [self.sig_node.index, self.int_node.index, self.gnd_node.index, self.sources[0].index]

    def stamp_params(self):
        return {"Z0": self.Z0, "port": self.index}


class _TwoTerminal(Element):
    _pname = "p"

    def stamp_ids(self):
        return [self.nodes[0].index, self.nodes[1].index]


class Resistor(_TwoTerminal):
    """g = 1/R with R a plain number (base/resistor.py:23-61)."""
    kind = "resistor"
    component_name = "Resistor"
    unit = "Ohm"

    def __init__(self, node1, node2, resistance):
        super().__init__([node1, node2])
        self.G: SimParam = enforce_simparam(lambda f: 1 / resistance)
        self.R: SimParam = enforce_simparam(resistance)
        self.value = resistance

    def stamp_params(self):
        return {"G": self.G}


class Admittance(_TwoTerminal):
    kind = "admittance"
    component_name = "Admittance"
    unit = "S"

    def __init__(self, node1, node2, admittance):
        super().__init__([node1, node2])
        self.Y: SimParam = enforce_simparam(admittance)
        self.value = self.Y(self.PRINT_FREQ)

    def stamp_params(self):
        return {"Y": self.Y}


class Impedance(_TwoTerminal):
    kind = "impedance"
    component_name = "Impedance"
    unit = "Ohm"

    def __init__(self, node1, node2, impedance):
        super().__init__([node1, node2])
        self.Z: SimParam = enforce_simparam(impedance)
        self.value = self.Z(self.PRINT_FREQ)

    def stamp_params(self):
        return {"Z": self.Z}


class Capacitor(_TwoTerminal):
    kind = "capacitor"
    component_name = "Capacitor"
    unit = "F"

    def __init__(self, node1, node2, capacitance):
        super().__init__([node1, node2])
        self.C: SimParam = enforce_simparam(capacitance)
        self.value = capacitance

    def stamp_params(self):
        return {"C": self.C}


class Inductor(_TwoTerminal):
    kind = "inductor"
    component_name = "Inductor"
    unit = "H"

    def __init__(self, node1, node2, inductance):
        super().__init__([node1, node2])
        self.L: SimParam = enforce_simparam(inductance)
        self.value = inductance

    def stamp_params(self):
        return {"L": self.L}


class TransmissionLine(Element):
    """Ideal line from ABCD parameters; ids = [n1, n2, gnd] (base/tl.py:5-67)."""
    kind = "tline"
    component_name = "TransmissionLine"
    unit = "Ohm"

    def __init__(self, node1, node2, gnd, Z0, length, beta):
        super().__init__([node1, node2, gnd])
        self.gnd = gnd
        self.nterminals = 4
        self.Z0: SimParam = enforce_simparam(Z0)
        self.beta: SimParam = enforce_simparam(beta)
        self.length: SimParam = enforce_simparam(length)
        self.value = self.Z0.value

    def stamp_ids(self):
        return [self.nodes[0].index, self.nodes[1].index, self.gnd.index]

    def stamp_params(self):
        return {"Z0": self.Z0, "beta": self.beta, "length": self.length}

    def record(self) -> dict:
        rec = super().record()
        rec["length"] = self.length.value          # what base/tl.py:52 reads at every run
        return rec


class NPortS(Element):
    """N-port given by S_ij(f) callables and a numeric reference impedance; ids = nodes + [gnd]
    (base/nport.py:5-45)."""
    kind = "nport_s"
    component_name = "NPort(S)"
    unit = "Ohm"

    def __init__(self, nodes, gnd, functions, Z0):
        super().__init__(nodes)
        self.gnd = gnd
        self.functions = functions
        self.Z0 = Z0
        self.value = self.Z0

    def stamp_ids(self):
        return [n.index for n in self.nodes] + [self.gnd.index]

    def stamp_params(self):
        return {"S": self.functions, "Z0": self.Z0}


class NPortY(Element):
    """N-port given by Y_ij(f) callables.  The reference's class cannot be constructed
    (base/nport.py:61 reads an attribute that is never set); this one works
    and stamps the (P+1)x(P+1) indefinite block the reference's compiler would have built."""
    kind = "nport_y"
    component_name = "NPort(Y)"
    unit = "Ohm"

    def __init__(self, nodes, gnd, functions, Z0):
        super().__init__(nodes)
        self.gnd = gnd
        self.functions = functions
        self.Z0: SimParam = enforce_simparam(Z0)
        self.nports = len(nodes)
        self.value = self.Z0.value

    def stamp_ids(self):
        return [n.index for n in self.nodes] + [self.gnd.index]

    def stamp_params(self):
        return {"Y": self.functions}
