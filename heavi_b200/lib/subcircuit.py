"""Sub-circuit parts: objects that collect terminal nodes and, once every terminal is attached,
expand into primitive components of the network they were attached to.

Behaviour matches the reference's part framework (heavi/lib/libgen.py:24-187) as user code sees it:

* `part.connect(n1, n2, ..., gnd=None)` attaches all terminals at once and expands immediately;
* `part.partial_connect(i, node)` / `part.t(i) > node` / `part.t(i) < node` attach one terminal and
  expand when terminals 1..n_nodes are all present;
* for two-terminal parts `part < a` attaches terminal 1 and `part > b` attaches terminal 2 and
  returns `b`, so parts chain as `a > part > b`... in the reference's notation;
* `part.node(i)` gives terminal i, creating a fresh network node on first use;
* subclasses set `n_nodes` (or override `__post_init__`) and implement `__on_connect__`.
"""
from __future__ import annotations

from ..netlist import Network, Node


class SubCircuit:
    def __init__(self):
        self.network: Network | None = None
        self.nodes: dict[int, Node] = {}
        self.gnd: Node | None = None
        self.n_nodes: int = 0
        self.__post_init__()
        self._selected_terminal: int | None = None

    # ---- hooks for subclasses -----------------------------------------------------------------
    def __post_init__(self):
        """Called from the constructor; fixed-arity parts set `n_nodes` here."""

    def __on_connect__(self):
        raise NotImplementedError("This method must be implemented in the child class.")

    # ---- terminal bookkeeping -------------------------------------------------------------------
    @property
    def first_node(self) -> int:
        """Lowest terminal number that has not been attached yet."""
        return next((i for i in range(1, self.n_nodes) if i not in self.nodes), self.n_nodes)

    def _adopt_network(self):
        for node in self.nodes.values():
            self.network = node._parent

    # the reference exposes these two under dunder names; keep them callable for subclasses
    __update__ = _adopt_network

    def __validate__(self):
        problems = []
        if self.n_nodes == 0:
            problems.append("The component must have at least one node.")
        elif not isinstance(self.network, Network):
            problems.append("The component must be connected to a Network object.")
        else:
            for node in self.nodes.values():
                if not isinstance(node, Node):
                    problems.append(f"All nodes must be Node objects. Got {type(node)} instead.")
                elif node._parent is not self.network:
                    problems.append("All nodes must belong to the same Network object.")
            if self.gnd is not None:
                if not isinstance(self.gnd, Node):
                    problems.append("The ground must be a Node object.")
                elif self.gnd._parent is not self.network:
                    problems.append("The ground node must belong to the same Network object.")
        if problems:
            raise ValueError(problems[0])

    def _complete(self) -> bool:
        return all(i in self.nodes for i in range(1, self.n_nodes + 1))

    def _check_full(self) -> None:
        if self._complete():
            self.__on_connect__()

    def node(self, index: int) -> Node:
        if index not in self.nodes:
            self.nodes[index] = self.network.node()
        return self.nodes[index]

    # ---- attaching ----------------------------------------------------------------------------------
    def partial_connect(self, index: int, node: Node) -> "SubCircuit":
        self.nodes[index] = node
        self._adopt_network()
        self._check_full()
        return self

    def connect(self, *nodes, gnd: Node = None) -> "SubCircuit":
        self.gnd = gnd
        self.nodes.update({i: n for i, n in enumerate(nodes, start=1)})
        self._adopt_network()
        if self.gnd is None:
            self.gnd = self.network.gnd
        self.__validate__()
        self.__on_connect__()
        return self

    def t(self, index: int) -> "SubCircuit":
        """Select terminal `index` for the next `<` / `>`."""
        self._selected_terminal = index
        return self

    def _attach_selected(self, other):
        if not isinstance(other, Node):
            return NotImplemented
        return self.partial_connect(self._selected_terminal, other)

    __lt__ = _attach_selected
    __gt__ = _attach_selected


class TwoNodeSubCircuit(SubCircuit):
    def __post_init__(self):
        self.n_nodes = 2

    def __lt__(self, other):
        if not isinstance(other, Node):
            return NotImplemented
        self.partial_connect(1, other)
        return self

    def __gt__(self, other):
        if not isinstance(other, Node):
            return NotImplemented
        self.partial_connect(2, other)
        return other


def _fixed_arity(n: int, name: str):
    def post_init(self):
        self.n_nodes = n
    return type(name, (SubCircuit,), {"__post_init__": post_init, "__doc__": f"Part with {n} terminals."})


ThreeNodeSubCircuit = _fixed_arity(3, "ThreeNodeSubCircuit")
FourNodeSubCircuit = _fixed_arity(4, "FourNodeSubCircuit")
