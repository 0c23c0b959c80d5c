"""SubCircuit framework: collect terminal nodes, then let the part add primitives to the network.

Behavioural equivalent of the reference's `heavi/lib/libgen.py:24-187`: `.connect(*nodes, gnd=)`,
`.partial_connect`, `.t(i) > node` / `.t(i) < node` terminal selection, `.node(i)` lazy internal
nodes, and the `__on_connect__` hook a part overrides.
"""
from __future__ import annotations

from ..netlist import Network, Node


class SubCircuit:
    def __init__(self):
        self.network: Network = None
        self.nodes: dict[int, Node] = {}
        self.gnd: Node = None
        self.n_nodes: int = 0
        self.__post_init__()
        self._selected_terminal: int = None

    def __post_init__(self):
        pass

    @property
    def first_node(self) -> int:
        for i in range(1, self.n_nodes):
            if i not in self.nodes:
                return i
        return self.n_nodes

    def __update__(self):
        for node in self.nodes.values():
            self.network = node._parent

    def __validate__(self):
        if self.n_nodes == 0:
            raise ValueError("The component must have at least one node.")
        if not isinstance(self.network, Network):
            raise ValueError("The component must be connected to a Network object.")
        for node in self.nodes.values():
            if not isinstance(node, Node):
                raise ValueError(f"All nodes must be Node objects. Got {type(node)} instead.")
            if node._parent is not self.network:
                raise ValueError("All nodes must belong to the same Network object.")
        if self.gnd is not None:
            if not isinstance(self.gnd, Node):
                raise ValueError("The ground must be a Node object.")
            if self.gnd._parent is not self.network:
                raise ValueError("The ground node must belong to the same Network object.")

    def _check_full(self) -> None:
        if all(i in self.nodes for i in range(1, self.n_nodes + 1)):
            self.__on_connect__()

    def __on_connect__(self):
        raise NotImplementedError("This method must be implemented in the child class.")

    def t(self, index: int) -> "SubCircuit":
        self._selected_terminal = index
        return self

    def __lt__(self, other):
        if isinstance(other, Node):
            self.partial_connect(self._selected_terminal, other)
            return self
        return NotImplemented

    __gt__ = __lt__

    def node(self, index: int) -> Node:
        if index not in self.nodes:
            self.nodes[index] = self.network.node()
        return self.nodes.get(index, None)

    def partial_connect(self, index: int, node: Node) -> "SubCircuit":
        self.nodes[index] = node
        self.__update__()
        self._check_full()
        return self

    def connect(self, *nodes, gnd: Node = None) -> "SubCircuit":
        self.gnd = gnd
        for i, node in enumerate(nodes):
            self.nodes[i + 1] = node
        self.__update__()
        if self.gnd is None:
            self.gnd = self.network.gnd
        self.__validate__()
        self.__on_connect__()
        return self


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


class ThreeNodeSubCircuit(SubCircuit):
    def __post_init__(self):
        self.n_nodes = 3


class FourNodeSubCircuit(SubCircuit):
    def __post_init__(self):
        self.n_nodes = 4
