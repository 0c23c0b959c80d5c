"""Netlist core: nodes, index allocation and the `Network` building API with `run_sp`.

Drop-in for the reference's `heavi/network.py` + `heavi/node.py` on the S-parameter path:
same builder signatures, same node/source numbering (`_define_indices`, network.py:109-127),
same `run_sp(frequencies, reinitialize=False) -> Sparameters` contract (network.py:354-421).
What differs is *how* `run_sp` computes: the netlist is frozen into a stamp table (`plan.py`) and
handed through the C-ABI (`engine.py` -> `libheavi_b200.so`) to sm_100a CUDA kernels that assemble,
factor and convert every (sample, frequency) point on the GPU.  There is no CPU fallback: without
the CUDA library `run_sp` raises.
"""
from __future__ import annotations

from typing import Callable

import numpy as np

from .elements import (Admittance, Capacitor, Element, Impedance, Inductor, MatSource, NPortS, NPortY,
                       Port, Resistor, TransmissionLine)
from .params import PhaseConstant, SimParam, enforce_simparam
from .results import DeviceStatSparameters, Sparameters, StatSparameters

try:                                    # same logger the reference uses; optional here
    from loguru import logger
except Exception:                       # pragma: no cover
    import logging
    logger = logging.getLogger("heavi_b200")


class _Unassigned:
    def __repr__(self):
        return "_"

    __str__ = __repr__


UNASSIGNED = _Unassigned()


class Node:
    """Circuit node (node.py:42-106).  `a > b` merges a into b and returns b."""

    def __init__(self, name: str, _index=UNASSIGNED, _parent=None, _linked: "Node" = None, _gnd: bool = False):
        self.name = name
        self._index = _index
        self._parent = _parent
        self._linked = _linked
        self._gnd = _gnd

    def __repr__(self) -> str:
        if self._gnd:
            return "Node_GND"
        if self._linked is None:
            return f"{self.name}[{self._index}]"
        return f"LinkedNode[{self._linked}]"

    __str__ = __repr__

    def set_index(self, index: int):
        self._index = index

    def unique(self) -> "Node":
        return self._linked if self._linked is not None else self

    @property
    def index(self) -> int:
        return self._linked.index if self._linked is not None else self._index

    def merge(self, other: "Node") -> "Node":
        self._linked = other
        return self

    def __gt__(self, other):
        if isinstance(other, Node):
            self._linked = other
            return other
        return NotImplemented


class NodeAllocator:
    """Named pool of numbered nodes (node.py:119-169)."""

    def __init__(self, index: str, parent):
        self._index = index
        self._pointer = 0
        self._nodes: dict[int, Node] = {}
        self._parent = parent

    def _next_free(self) -> int:
        number = self._pointer
        self._pointer += 1
        while self._pointer in self._nodes:
            self._pointer += 1
        return number

    def request(self, number: int = None):
        if number is None:
            number = self._next_free()
        created = number not in self._nodes
        if created:
            self._nodes[number] = Node(f"{self._index}{number}", _parent=self._parent)
        return self._nodes[number], created


class Network:
    """Netlist container and S-parameter entry point (network.py:63-702)."""

    def __init__(self, default_name: str = "Node", suppress_loadbar: bool = False):
        self.gnd = Node("gnd", _parent=self, _gnd=True)
        self.components: list[Element] = []
        self.sources: list[MatSource] = []
        self.ports: dict[int, Port] = {}
        self.port_counter = 1
        self.suppress_loadbar = suppress_loadbar      # accepted for compatibility; there is no load bar
        self._default_node_index = default_name
        self._nodes: list[Node] = [self.gnd]
        self._node_library: dict[str, NodeAllocator] = {default_name: NodeAllocator(default_name, self)}
        self._initialized = False
        self._engine_cache = None                      # (netlist signature, CompiledNetwork)

    # ------------------------------------------------------------------ bookkeeping
    def print_components(self) -> None:
        for comp in self.components:
            print(comp.summary())

    @property
    def node_names(self) -> list[str]:
        return [n.name for n in self._nodes]

    def unlinked_nodes(self) -> list[Node]:
        return [n for n in self._nodes if n._linked is None]

    def port(self, number: int):
        return self.ports.get(number, None)

    def get_port_index(self) -> int:
        index = self.port_counter
        self.port_counter += 1
        return index

    def node(self, number: int = None, index: str = None) -> Node:
        if index is None:
            index = self._default_node_index
        if index not in self._node_library:
            self._node_library[index] = NodeAllocator(index, self)
        node, is_new = self._node_library[index].request(number)
        if is_new:
            self._nodes.append(node)
        return node

    def mnodes(self, N: int, name: str = None) -> list[Node]:
        if name is None:
            return [self.node() for _ in range(N)]
        return [self.node(f"{name}_{i + 1}") for i in range(N)]

    # ------------------------------------------------------------------ index allocation
    def _define_indices(self) -> None:
        """Node index = position among unmerged nodes in creation order (ground first = 0);
        source index continues after the last node, one per port in component order
        (network.py:109-127).  Unlike the reference, calling it again does not grow
        `self.sources`."""
        i = 0
        for node in self._nodes:
            if node._linked is not None:
                continue
            node.set_index(i)
            i += 1
        self.sources = []
        for comp in self.components:
            srcs = [MatSource(k) for k in range(i, i + comp.sp_sources)]
            self.sources.extend(srcs)
            comp.set_sources(srcs)
            i += comp.sp_sources

    def _check_unconnected_nodes(self) -> None:
        used = {id(n.unique()) for comp in self.components for n in comp.all_nodes}
        for node in self._nodes:
            u = node.unique()
            if id(u) not in used:
                logger.error(f"Node {u.name} is not connected to any components.")
                logger.error("Unconnected nodes will cause the analysis to yield 0 values.")
                raise ValueError(f"Node {u.name} is not connected to any components.")

    def _prepare(self) -> None:
        if not self._initialized:
            self._define_indices()
            self._check_unconnected_nodes()
            self._initialized = True

    @property
    def n_nodes(self) -> int:
        return max(n.index for n in self._nodes) + 1

    # ------------------------------------------------------------------ stamp table / engine
    def stamp_table(self):
        """Freeze the netlist: see `plan.StampTable`."""
        from .plan import StampTable
        self._prepare()
        return StampTable.from_network(self)

    def records(self) -> list[dict]:
        """Neutral per-component records (kind, ids, parameters) for the test oracle."""
        self._prepare()
        return [c.record() for c in self.components]

    def _compiled(self):
        from .engine import CompiledNetwork
        self._prepare()
        sig = (len(self.components), len(self._nodes), tuple(id(c) for c in self.components),
               tuple(n.index for n in self._nodes))
        if self._engine_cache is None or self._engine_cache[0] != sig:
            self._engine_cache = (sig, CompiledNetwork(self))
        return self._engine_cache[1]

    # ------------------------------------------------------------------ analyses
    def run_sp(self, frequencies: np.ndarray, reinitialize: bool = False) -> Sparameters:
        """S-parameter sweep at `frequencies` -> Sparameters(S[P,P,nF] complex128, f float32)."""
        compiled = self._compiled()
        f64 = np.ascontiguousarray(frequencies, dtype=np.float64)
        # float32 copy of the frequency axis, the reference's quirk (network.py:407,421); float64 holds
        # every float32-representable input exactly, so converting from f64 gives the same values.  The
        # kernels produce it and it comes back with the tiles of S.
        from .engine import pinned_empty
        f32 = pinned_empty(f64.shape, np.float32) if f64.ndim == 1 else None
        S = compiled.run(f64, f32_out=f32 if f64.ndim == 1 else None)   # [1,P,P,nF] on pinned host memory
        if f64.ndim != 1:
            f32 = f64.astype(np.float32)
        if reinitialize:
            self._initialized = False
        return Sparameters(S[0], f32)

    def run_sp_async(self, frequencies: np.ndarray):
        """`run_sp` that returns at once: the sweep is queued on the GPU and `.result()` of the returned object
        gives the `Sparameters`.  Independent models overlap -- `run_sp_many` is the convenience form."""
        compiled = self._compiled()
        f64 = np.ascontiguousarray(frequencies, dtype=np.float64)
        from .engine import pinned_empty
        f32 = pinned_empty(f64.shape, np.float32)
        return PendingSparameters(compiled.run_async(f64, f32_out=f32), f32)

    def run_sp_batch(self, frequencies: np.ndarray, drivers, values: np.ndarray) -> np.ndarray:
        """Batched Monte-Carlo / sweep evaluation: `values[s, k]` is the value of `drivers[k]`
        (Random / Param objects) in sample s.  Returns S[nS,P,P,nF]; every (sample, frequency)
        point is solved in one GPU pass instead of one `run_sp` call per sample."""
        compiled = self._compiled()
        f64 = np.ascontiguousarray(np.asarray(frequencies, dtype=np.float64))
        return compiled.run(f64, drivers=list(drivers), values=np.asarray(values, dtype=np.float64))

    def run_sp_resident(self, frequencies, drivers=None, values=None):
        """`run_sp` / `run_sp_batch` whose result stays on the GPU: torch CUDA tensor S[nS,P,P,nF]
        (complex128, nS = 1 for a plain sweep).  Use it when the consumer of S is GPU code."""
        return self._compiled().run_resident(frequencies, None if drivers is None else list(drivers),
                                             None if values is None else np.asarray(values, dtype=np.float64))

    def run_mc(self, frequencies: np.ndarray, mc, N: int) -> StatSparameters:
        """`for i in mc.iterate(N): mc.submit_S(run_sp(f))` (numeric.py:498-507) as one batch.
        Draws are made on the host in the reference's order, so results are seed-compatible."""
        values = mc.draw(N)
        S4 = self.run_sp_batch(frequencies, mc._random_numbers, values)
        for k, r in enumerate(mc._random_numbers):     # leave the objects as the serial loop would
            r._value = values[-1, k] if N > 0 else r._value
        mc.S = StatSparameters(N)
        mc.S.add_batch(S4)
        return mc.S

    def run_mc_stats(self, frequencies: np.ndarray, mc, N: int) -> DeviceStatSparameters:
        """Monte-Carlo run whose statistics are reduced on the GPU: same draws as `run_mc`, but only
        mean / std of |S|, |S|^2 and sqrt(|S|^2) per (port pair, frequency) come back over PCIe."""
        values = mc.draw(N)
        compiled = self._compiled()
        f64 = np.ascontiguousarray(np.asarray(frequencies, dtype=np.float64))
        moments = compiled.run_stats(f64, mc._random_numbers, values)
        for k, r in enumerate(mc._random_numbers):
            r._value = values[-1, k] if N > 0 else r._value
        return DeviceStatSparameters(moments, N, np.array(frequencies).astype(np.float32))

    def run_sweep(self, frequencies: np.ndarray, sweep):
        """`for ix, vals in sweep.iterate(): sweep.submit(run_sp(f), f)` (numeric.py:418-446, :386-390)
        as one batch: every grid point of the ParameterSweep is one sample.  Returns `sweep.S`
        (an NDSparameters of shape dims + [P, P, nF]) with the same contents the serial loop builds."""
        from .results import NDSparameters
        combos = sweep.combinations()
        params = sweep.params
        values = np.array([[p._values[i] for i, dim in zip(ixs, sweep.sweep_dimensions) for p in dim]
                           for ixs in combos], dtype=np.float64).reshape(len(combos), len(params))
        S4 = self.run_sp_batch(frequencies, params, values)
        shape = [len(dim[0]) for dim in sweep.sweep_dimensions]
        ports = np.arange(1, S4.shape[1] + 1)
        f32 = np.array(frequencies).astype(np.float32)
        axes = [np.array([p._values for p in dim]) for dim in sweep.sweep_dimensions]
        sweep._fdata = f32
        sweep._mdim_data = NDSparameters(np.ascontiguousarray(S4).reshape(shape + list(S4.shape[1:])),
                                         axes + [ports, ports, f32])
        for p, v in zip(params, values[-1] if len(values) else []):       # state after the serial loop
            p._value = v
        return sweep._mdim_data

    def run_dc(self, *a, **k):
        raise NotImplementedError("DC analysis is outside the S-parameter hot path this package implements")

    def run_transient(self, *a, **k):
        raise NotImplementedError("transient analysis is outside the S-parameter hot path this package implements")

    # ------------------------------------------------------------------ builders
    def _add(self, comp: Element) -> Element:
        self.components.append(comp)
        self._initialized = False        # indices are recomputed (deterministically) on the next run
        return comp

    def new_port(self, Z0: float, node: Node = None, gnd: Node = None) -> Port:
        """network.py:464-488: the internal node is created before the signal node."""
        if gnd is None:
            gnd = self.gnd
        int_node = self.node()
        sig = self.node() if node is None else node
        port = Port(self.get_port_index(), sig, int_node, gnd, Z0)
        self._add(port)
        self.ports[port.index] = port
        return port

    def quick_port(self, Z0: float, gnd: Node = None) -> Node:
        # the reference forwards `gnd` as the *signal node* argument (network.py:500); kept
        return self.new_port(Z0, gnd).sig_node

    def admittance(self, node1: Node, node2: Node, Y) -> Admittance:
        return self._add(Admittance(node1, node2, Y))

    def impedance(self, node1: Node, node2: Node, Z, display_value: float = None) -> Impedance:
        return self._add(Impedance(node1, node2, Z))

    def resistor(self, node1: Node, node2: Node, R: float) -> Resistor:
        return self._add(Resistor(node1, node2, R))

    def capacitor(self, node1: Node, node2: Node, C) -> Capacitor:
        return self._add(Capacitor(node1, node2, C))

    def inductor(self, node1: Node, node2: Node, L) -> Inductor:
        return self._add(Inductor(node1, node2, L))

    def transmissionline(self, port1: Node, port2: Node, Z0, er, L: float, gnd: Node = None) -> TransmissionLine:
        """beta(f) = 2 pi f / c0 sqrt(er(f)) (network.py:608-633)."""
        return self.TL(port1, port2, PhaseConstant(enforce_simparam(er)), L, enforce_simparam(Z0, unit="Ω"), gnd)

    def TL(self, node1: Node, node2: Node, beta, length: float, Z0, gnd: Node = None) -> TransmissionLine:
        if gnd is None:
            gnd = self.gnd
        return self._add(TransmissionLine(node1, node2, gnd, Z0, length, beta))

    def n_port_S(self, gnd: Node, nodes: list[Node], Sparam: list[list[Callable]], Z0: float) -> NPortS:
        return self._add(NPortS(nodes, gnd, Sparam, Z0))

    def n_port_Y(self, gnd: Node, nodes: list[Node], Yparams: list[list[Callable]], Z0: float) -> NPortY:
        return self._add(NPortY(nodes, gnd, Yparams, Z0))

    def diode(self, *a, **k):
        raise NotImplementedError("diodes belong to the DC/transient analyses, outside this package's scope")

    DC_source = TRANSIENT_source = diode


class PendingSparameters:
    """What `Network.run_sp_async` returns: `.result()` waits for the sweep and wraps it like `run_sp` does."""

    __slots__ = ("_pending", "_f32")

    def __init__(self, pending, f32):
        self._pending, self._f32 = pending, f32

    def result(self) -> Sparameters:
        return Sparameters(self._pending.result()[0], self._f32)


def run_sp_many(jobs) -> list:
    """[(network, frequencies), ...] -> [Sparameters, ...]: every sweep is queued before the first result is
    awaited, so independent models (a filter set, the candidates of a design study) share the GPU and the PCIe
    link back to back instead of paying each call's fixed costs one after the other."""
    pending = [net.run_sp_async(f) for net, f in jobs]
    return [p.result() for p in pending]
