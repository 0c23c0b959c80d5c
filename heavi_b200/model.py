"""`Model` = `Network` + the filter library (reference: heavi/model.py:26-49)."""
from __future__ import annotations

from .filters import Filtering
from .netlist import Network, Node


class Model(Network):
    def __init__(self, default_name: str = "Node", filter_library=Filtering, suppress_loadbar: bool = False):
        super().__init__(default_name, suppress_loadbar=suppress_loadbar)
        self.filters = filter_library(self)
        self.numbered_nodes: dict[int, Node] = {}

    def __call__(self, index: int) -> Node:
        """`model(3)` -> the node remembered under number 3, created on first use."""
        if index not in self.numbered_nodes:
            self.numbered_nodes[index] = self.node()
        node = self.numbered_nodes[index]
        node._parent = self
        return node
