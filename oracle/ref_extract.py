"""Turn a *reference* `heavi.Network` (container only) into the oracle's neutral records, and
capture the reference's own intermediate results (dense MNA tensor, port table, S).

TEST INFRASTRUCTURE ONLY (used by oracle/make_golden.py and oracle/check_oracle_vs_ref.py).
"""
from __future__ import annotations

import numpy as np


def initialise(net):
    """The first-call block of Network.run_sp (network.py:372-376)."""
    from heavi.base import SimulationType
    if not net._initialized:
        net._assign_internal_nodes(SimulationType.SP)
        net._define_indices(SimulationType.SP)
        net._check_unconnected_nodes()
        net._initialized = True


def sizes(net):
    M = len(net.sources)
    N = max(node.index for node in net._nodes) + 1
    return N, M


def records_from_reference(net):
    initialise(net)
    recs = []
    for c in net.components:
        name = type(c).__name__
        if name == "Port":
            recs.append({"kind": "port", "port": c.index, "Z0": c.Z0,
                         "ids": [c.sig_node.index, c.int_node.index, c.gnd_node.index, c.sources[0].index]})
        elif name == "Resistor":
            recs.append({"kind": "resistor", "ids": [n.index for n in c.nodes], "G": c.G})
        elif name == "Admittance":
            recs.append({"kind": "admittance", "ids": [n.index for n in c.nodes], "Y": c.Y})
        elif name == "Impedance":
            recs.append({"kind": "impedance", "ids": [n.index for n in c.nodes], "Z": c.Z})
        elif name == "Capacitor":
            recs.append({"kind": "capacitor", "ids": [n.index for n in c.nodes], "C": c.C})
        elif name == "Inductor":
            recs.append({"kind": "inductor", "ids": [n.index for n in c.nodes], "L": c.L})
        elif name == "TransmissionLine":
            recs.append({"kind": "tline", "ids": [c.nodes[0].index, c.nodes[1].index, c.gnd.index],
                         "Z0": c.Z0, "beta": c.beta, "length": c.length.value})
        elif name == "NPortS":
            recs.append({"kind": "nport_s", "ids": [n.index for n in c.nodes] + [c.gnd.index],
                         "S": c.functions, "Z0": c.Z0})
        else:
            raise NotImplementedError(name)
    return recs


def reference_assemble(net, f):
    """The assembly half of Network.run_sp, network.py:380-404, returning what it builds."""
    from heavi.base import SimulationType
    initialise(net)
    N, M = sizes(net)
    nF = len(f)
    compilers = [c.SP_matrix_compiler() for c in net._get_components(SimulationType.SP)]
    mna = np.zeros((M + N, M + N, nF), dtype=np.complex128)
    for comp in compilers:
        if comp is not None:
            mna = comp(mna, f)
    Zs = np.zeros((M, nF), dtype=np.complex128)
    rows = []
    for iport, port in net.ports.items():
        rows.append((iport - 1, port.sig_node.index, port.int_node.index, port.gnd_node.index))
        Zs[iport - 1, :] = port.Z0(f)
    return mna, Zs, np.array(rows).astype(np.int32), N, M


def flat_ids(recs):
    """Component kinds and ids as flat arrays (what the golden files store)."""
    kinds = np.array([r["kind"] for r in recs])
    off = np.cumsum([0] + [len(r["ids"]) for r in recs]).astype(np.int32)
    ids = np.array([i for r in recs for i in r["ids"]], dtype=np.int32)
    return kinds, off, ids
