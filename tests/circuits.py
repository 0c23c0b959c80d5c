"""Circuit builders written against the *public building API* only.

Every builder takes the package object `hv` -- either the real reference (`heavi`, through
`oracle/ref_shim.py`, container only) or the product front-end (`heavi_b200`) -- and returns
`(model, f)` (plus a Monte-Carlo driver where one exists).  Because the same code drives both,
a golden vector generated from the reference pins node/source indexing, stamp order and S for
the product.

BASELINE.json configs:  c1 README demo, c2 demo2 Cauer filter set, c3 synthetic 8-port
Touchstone block, c4 Monte-Carlo Cauer band-pass, c5 ladder/TL network.
"""
from __future__ import annotations

import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SYNTH8 = os.path.join(HERE, "golden", "synth8.snp")


def _model(hv):
    return hv.Model(suppress_loadbar=True)


# ---------------------------------------------------------------- config 1 (README.md:34-52)
def build_c1(hv, nF: int = 2001):
    smd = hv.lib.smd
    m = _model(hv)
    n1 = m.node()
    m.new_port(50, n1)
    n2 = m.node()
    n3 = m.node()
    m.new_port(50, n3)
    smd.SMDResistor(5, smd.SMDResistorSize.R0402).connect(n2, n3)
    m.filters.cauer_filter(m.gnd, n1, n2, 2e9, 70e6, 5, 0.03, hv.FilterType.CHEBYCHEV,
                           type=hv.BandType.BANDPASS)
    return m, hv.frange(1.8e9, 2.2e9, nF)


# ---------------------------------------------------------------- config 2 (examples/demo2:13-24)
def build_c2(hv, band: str = "bandpass", nF: int = 100_001):
    m = _model(hv)
    n1 = m.quick_port(50)
    n2 = m.quick_port(50)
    bt = {"bandpass": hv.BandType.BANDPASS, "lowpass": hv.BandType.LOWPASS,
          "highpass": hv.BandType.HIGHPASS, "bandstop": hv.BandType.BANDSTOP}[band]
    m.filters.cauer_filter(m.gnd, n1, n2, 2.5e9, 0.2e9, 5, 0.05, hv.FilterType.CHEBYCHEV, bt, 50)
    return m, hv.frange(2e9, 3e9, nF)


# ---------------------------------------------------------------- config 3
def synth8_data(nfs: int = 201):
    """Reciprocal passive synthetic 8-port: S(f) = (Q diag(g e^{-j 2 pi f tau})) Q^T."""
    rng = np.random.default_rng(1234)
    fs = np.linspace(0.5e9, 6.5e9, nfs)
    Q, _ = np.linalg.qr(rng.normal(size=(8, 8)) + 1j * rng.normal(size=(8, 8)))
    tau = rng.uniform(0.05e-9, 0.4e-9, 8)
    g = rng.uniform(0.2, 0.7, 8)
    S = np.empty((nfs, 8, 8), dtype=complex)
    for k, fk in enumerate(fs):
        S[k] = (Q * (g * np.exp(-2j * np.pi * fk * tau))) @ Q.T
    return fs, S


def write_synth8(path: str = SYNTH8) -> str:
    """Touchstone v1, `# hz s ri r 50`, one line per frequency, single-space separated
    (the reference's parser splits on single spaces only, touchstone.py:204)."""
    fs, S = synth8_data()
    with open(path, "w") as fh:
        fh.write("! synthetic reciprocal passive 8-port, default_rng(1234) (tests/circuits.py)\n")
        fh.write("# hz s ri r 50\n")
        for k, fk in enumerate(fs):
            nums = ["%.17g" % fk]
            for v in S[k].reshape(-1):
                nums.append("%.17g" % v.real)
                nums.append("%.17g" % v.imag)
            fh.write(" ".join(nums) + "\n")
    return path


def build_c3(hv, nF: int = 1_000_000, path: str = SYNTH8):
    m = _model(hv)
    blk = hv.FileBasedNPort(path)
    inner = [m.node() for _ in range(8)]
    for k, n in enumerate(inner):
        sig = m.quick_port(50)
        mid = m.node()
        m.inductor(sig, mid, (1.0 + 0.25 * k) * 1e-9)
        m.capacitor(mid, m.gnd, (0.5 + 0.1 * k) * 1e-12)
        m.capacitor(mid, n, (2.0 + 0.3 * k) * 1e-12)
    blk.connect(*inner)
    return m, np.linspace(1e9, 6e9, nF)


# ---------------------------------------------------------------- config 4
def build_c4(hv, nF: int = 2001):
    """c1's Cauer band-pass (no SMD part) with every L and C ~ N(v, (0.02 v / 3)^2)."""
    nominal = _model(hv)
    a = nominal.quick_port(50)
    b = nominal.quick_port(50)
    caps, inds, _ = nominal.filters.cauer_filter(nominal.gnd, a, b, 2e9, 70e6, 5, 0.03,
                                                 hv.FilterType.CHEBYCHEV, type=hv.BandType.BANDPASS)
    values = {id(c): c.value for c in caps + inds}

    mc = hv.MonteCarlo()
    m = _model(hv)
    p1 = m.quick_port(50)
    p2 = m.quick_port(50)
    # same topology, rebuilt component by component in the nominal model's creation order
    node_map = {id(nominal.gnd): m.gnd, id(a): p1, id(b): p2}

    def mapped(node):
        key = id(node)
        if key not in node_map:
            node_map[key] = m.node()
        return node_map[key]

    for comp in nominal.components:
        if id(comp) not in values:
            continue
        v = values[id(comp)]
        rv = mc.gaussian(v, 0.02 * v / 3)
        n1, n2 = mapped(comp.nodes[0]), mapped(comp.nodes[1])
        if comp.component_name == "Capacitor":
            m.capacitor(n1, n2, rv)
        else:
            m.inductor(n1, n2, rv)
    return m, hv.frange(1.8e9, 2.2e9, nF), mc


def build_random_twoport(hv, nF: int = 41):
    """Monte-Carlo cable model (lib/mc.py RandomTwoPort) between a port and a short matched line: its
    S_ij are closures over `Random` objects, i.e. tabulated callables that change from trial to trial."""
    import importlib
    RandomTwoPort = importlib.import_module(hv.__name__ + ".lib.mc").RandomTwoPort
    m = _model(hv)
    mc = hv.MonteCarlo()
    a = m.quick_port(50)
    b = m.quick_port(50)
    mid = m.node()
    blk = RandomTwoPort(mc, 0.25, 50.0, VSWR=1.4, er=2.1, attenuation=0.8, phase_stability=4.0,
                        amplitude_stability=0.3, Z0_error=3.0)
    blk.connect(a, mid)
    m.transmissionline(mid, b, 50.0, 2.9, 7e-3)
    m.capacitor(mid, m.gnd, mc.gaussian(0.3e-12, 0.01e-12))
    return m, np.linspace(1e9, 4e9, nF), mc


# ---------------------------------------------------------------- config 5
def build_c5(hv, nF: int = 1_000_000, K: int = 190):
    rng = np.random.default_rng(2024)
    m = _model(hv)
    chain = [m.node() for _ in range(K + 1)]
    taps = [round(i * K / 7) for i in range(8)]
    for t in taps:
        m.new_port(50, chain[t])
    for k in range(K):
        if k % 2 == 0:
            m.inductor(chain[k], chain[k + 1], 10 ** rng.uniform(np.log10(0.5e-9), np.log10(5e-9)))
        else:
            m.transmissionline(chain[k], chain[k + 1], rng.uniform(30, 90), rng.uniform(2, 4),
                               rng.uniform(1e-3, 10e-3))
        if k % 5 == 0:
            m.resistor(chain[k], chain[k + 1], rng.uniform(200, 2000))
    for k in range(K + 1):
        m.capacitor(chain[k], m.gnd, 10 ** rng.uniform(np.log10(0.1e-12), np.log10(2e-12)))
        if k % 3 == 0:
            m.resistor(chain[k], m.gnd, rng.uniform(500, 5000))
    return m, np.linspace(0.5e9, 6e9, nF)


def build_band_ladder(hv, K: int = 60, reach: int = 2, nports: int = 3, seed: int = 0, nF: int = 40):
    """Chain of K+1 nodes whose elements couple node k with k+1..k+reach: the MNA matrix is banded with
    half bandwidth `reach` (after the plan compiler's ordering), the shape the banded kernel serves."""
    rng = np.random.default_rng(1000 + seed)
    m = _model(hv)
    chain = [m.node() for _ in range(K + 1)]
    for t in [round(i * K / max(nports - 1, 1)) for i in range(nports)]:
        m.new_port(float(rng.choice([25.0, 50.0, 75.0])), chain[t])
    for k in range(K):
        for d in range(1, reach + 1):
            if k + d > K:
                continue
            kind = rng.integers(0, 4)
            if d > 1 and rng.random() < 0.3:
                continue
            if kind == 0:
                m.inductor(chain[k], chain[k + d], 10 ** rng.uniform(-9.3, -8.3))
            elif kind == 1:
                m.resistor(chain[k], chain[k + d], rng.uniform(5, 500))
            elif kind == 2:
                m.capacitor(chain[k], chain[k + d], 10 ** rng.uniform(-12.5, -11.5))
            else:
                m.transmissionline(chain[k], chain[k + d], rng.uniform(30, 90), rng.uniform(2, 4), rng.uniform(1e-3, 10e-3))
    for k in range(K + 1):
        if k % 2 == 0:
            m.capacitor(chain[k], m.gnd, 10 ** rng.uniform(-13, -12))
        if k % 7 == 0:
            m.resistor(chain[k], m.gnd, rng.uniform(500, 5000))
    return m, np.linspace(0.5e9, 6e9, nF)


# ---------------------------------------------------------------- analytic known answers
def build_divider(hv):
    """10 ohm series resistor between a 50 ohm and a 75 ohm port."""
    m = _model(hv)
    a = m.quick_port(50)
    b = m.quick_port(75)
    m.resistor(a, b, 10.0)
    return m, np.array([1e6, 1e9, 3e9])


def build_matched_line(hv, Z0=50.0, er=2.2, length=0.0123):
    m = _model(hv)
    a = m.quick_port(Z0)
    b = m.quick_port(Z0)
    m.transmissionline(a, b, Z0, er, length)
    return m, np.linspace(0.2e9, 9e9, 64)


# ---------------------------------------------------------------- seeded random family
_KINDS = ("resistor", "capacitor", "inductor", "impedance", "admittance", "tline", "tline_lossy",
          "nport2", "impedance_fn", "self_loop")


def build_random(hv, seed: int):
    """Small random connected circuit exercising every stamp class and the quirks of App. A/B:
    callable parameters, lossy lines (complex er), an N-port S block with analytic callables,
    a port referenced to a non-ground node, a merged node and a same-node ("self loop") part."""
    rng = np.random.default_rng(10_000 + seed)
    m = _model(hv)
    n_nodes = int(rng.integers(2, 8))
    n_ports = int(rng.integers(1, 4))
    nodes = [m.node() for _ in range(n_nodes)]
    if seed % 5 == 3 and n_nodes > 3:
        extra = m.node()
        extra > nodes[1]              # merge: `extra` becomes an alias of nodes[1] (node.py:101-106)
        nodes.append(extra)
    port_nodes = list(rng.choice(n_nodes, size=min(n_ports, n_nodes), replace=False))
    for q, pn in enumerate(port_nodes):
        z0 = float(rng.choice([25.0, 50.0, 75.0, 100.0]))
        if seed % 7 == 2 and q == 0 and n_nodes > 2:
            ref = nodes[(pn + 1) % n_nodes]           # port against a non-ground reference node
            m.new_port(z0, nodes[pn], ref)
            m.resistor(ref, m.gnd, float(rng.uniform(5, 50)))
        else:
            m.new_port(z0, nodes[pn])

    def pick2():
        i, j = rng.choice(len(nodes) + 1, size=2, replace=False)
        pool = nodes + [m.gnd]
        return pool[i], pool[j]

    # spanning chain keeps every node attached to something
    for k in range(len(nodes) - 1):
        m.inductor(nodes[k], nodes[k + 1], float(10 ** rng.uniform(-9.5, -8)))
    for k in range(len(nodes)):
        m.capacitor(nodes[k], m.gnd, float(10 ** rng.uniform(-12.5, -11.5)))
    for _ in range(int(rng.integers(2, 9))):
        kind = _KINDS[int(rng.integers(0, len(_KINDS)))]
        a, b = pick2()
        if kind == "resistor":
            m.resistor(a, b, float(rng.uniform(5, 500)))
        elif kind == "capacitor":
            m.capacitor(a, b, float(10 ** rng.uniform(-13, -11)))
        elif kind == "inductor":
            m.inductor(a, b, float(10 ** rng.uniform(-10, -8)))
        elif kind == "impedance":
            m.impedance(a, b, complex(rng.uniform(5, 200), rng.uniform(-100, 100)))
        elif kind == "admittance":
            m.admittance(a, b, complex(rng.uniform(1e-3, 5e-2), rng.uniform(-2e-2, 2e-2)))
        elif kind == "tline":
            m.transmissionline(a, b, float(rng.uniform(30, 90)), float(rng.uniform(1, 5)),
                               float(rng.uniform(2e-3, 30e-3)))
        elif kind == "tline_lossy":
            m.transmissionline(a, b, float(rng.uniform(30, 90)),
                               complex(rng.uniform(2, 4), -rng.uniform(0.01, 0.1)),
                               float(rng.uniform(2e-3, 30e-3)))
        elif kind == "nport2":
            tau = float(rng.uniform(0.05e-9, 0.3e-9))
            r = float(rng.uniform(0.05, 0.4))
            t = float(rng.uniform(0.3, 0.8))
            s11 = lambda f, r=r, tau=tau: r * np.exp(-2j * np.pi * f * tau * 0.5)
            s21 = lambda f, t=t, tau=tau: t * np.exp(-2j * np.pi * f * tau)
            s12 = lambda f, t=t, tau=tau: 0.5 * t * np.exp(-2j * np.pi * f * tau)   # non-reciprocal
            s22 = lambda f, r=r: r * np.ones_like(f) + 0j
            m.n_port_S(m.gnd, [a, b], [[s11, s12], [s21, s22]], float(rng.choice([50.0, 75.0])))
        elif kind == "impedance_fn":
            R, L, C = float(rng.uniform(1, 100)), float(10 ** rng.uniform(-10, -9)), float(10 ** rng.uniform(-13, -12))
            m.impedance(a, b, lambda f, R=R, L=L, C=C: R + 2j * np.pi * f * L + 1 / (2j * np.pi * f * C))
        elif kind == "self_loop":
            # both terminals on one node: numpy `+=` keeps the last write only (App. A)
            m.resistor(a, a, float(rng.uniform(20, 200)))
    f = np.sort(rng.uniform(0.3e9, 8e9, size=int(rng.integers(3, 9))))
    return m, f


def build_twoport_cascade(hv, sections: int = 6, nF: int = 33, seed: int = 0):
    """Cascade of measured-style two-port S blocks (in-memory EMerge data), lines and shunt capacitors
    between two ports: chain-like, so narrow-banded, with an N-port (two-port) value op per section."""
    import importlib
    EMergeModel = importlib.import_module(hv.__name__ + ".lib.emerge").EMergeModel
    rng = np.random.default_rng(500 + seed)
    F = np.linspace(0.2e9, 8e9, 41)
    m = _model(hv)
    a, b = m.quick_port(50), m.quick_port(50)
    at = a
    for k in range(sections):
        # a slightly lossy, slightly mismatched line section as S data
        th = 2 * np.pi * F * (0.6e-10 + 0.2e-10 * rng.random())
        g = 0.05 + 0.1 * rng.random()
        t = (1 - g ** 2) ** 0.5 * 10 ** (-0.02) * np.exp(-1j * th)
        S = np.empty((len(F), 2, 2), dtype=complex)
        S[:, 0, 0] = g * np.exp(1j * 0.3 * k); S[:, 1, 1] = -g * np.exp(-1j * 0.2 * k)
        S[:, 0, 1] = t; S[:, 1, 0] = t
        n1, n2 = m.node(), m.node()
        m.transmissionline(at, n1, 50.0 + 5 * k, 2.2, 3e-3)
        EMergeModel((F, S)).connect(n1, n2)
        m.capacitor(n2, m.gnd, (0.1 + 0.05 * k) * 1e-12)
        at = n2
    m.inductor(at, b, 0.8e-9)
    return m, np.linspace(0.5e9, 6e9, nF)


# ---------------------------------------------------------------- optimiser / sweep callers
def build_optim(hv):
    """Small LC match driven by three optimiser variables and three band requirements
    (lib/optim.py: Optimiser.cap / ind / add_splimit / generate_objective, :239-438)."""
    import importlib
    optim = importlib.import_module(hv.__name__ + ".lib.optim")
    m = _model(hv)
    opt = optim.Optimiser(m)
    C1, L1, C2 = opt.cap(initial=-12.0), opt.ind(initial=-8.7), opt.cap(logscale=False, initial=-12.3)
    a, b = m.quick_port(50), m.quick_port(50)
    x = m.node()
    m.capacitor(a, m.gnd, C1)
    m.inductor(a, x, L1)
    m.capacitor(x, m.gnd, C2)
    m.resistor(x, b, 3.0)
    opt.add_splimit(1e9, 2e9, 21, (2, 1), lower_limit=optim.dBabove(-1))
    opt.add_splimit(4e9, 6e9, 31, (2, 1), upper_limit=optim.dBbelow(-20))
    opt.add_splimit(1e9, 2e9, 21, (1, 1), upper_limit=optim.dBbelow(-15), weight=0.5)
    return m, opt


def optim_population():
    rng = np.random.default_rng(4)
    return np.stack([rng.uniform(-12.5, -11.5, 9), rng.uniform(-9.2, -8.4, 9), rng.uniform(2e-13, 2e-12, 9)])   # [vars, candidates]


def build_sweep(hv):
    """4 x 3 ParameterSweep over a C and an L (numeric.py:314-449)."""
    m = _model(hv)
    sweep = hv.ParameterSweep()
    C = sweep.lin(0.5e-12, 2e-12).add(4)
    L = sweep.lin(1e-9, 3e-9).add(3)
    for p in (C, L):                       # the reference evaluates a component's value when it is created
        p.initialize()
    a, b = m.quick_port(50), m.quick_port(75)
    mid = m.node()
    m.inductor(a, mid, L)
    m.capacitor(mid, m.gnd, C)
    m.resistor(mid, b, 12.0)
    return m, sweep, np.linspace(1e9, 4e9, 257)


def build_hub(hv, n_block: int = 4, chain_rows=(2, 1, 0, 3), port_on=(0, 0, None, 1), seed: int = 0, nF: int = 33,
              scale: float = 1.0, diag_bias: float = 0.0):
    """An N-port S block (analytic passive S(f) given as Python callables) whose node k carries a chain of
    `chain_rows[k]` extra nodes; `port_on[k]` = position on that chain (0 = far end) of a port, None = no port,
    and for a chain of 0 rows a port value of 0 puts the port on the block node itself.  `scale` / `diag_bias`:
    S -> scale * S - diag_bias * I (still passive for scale * 0.7 + diag_bias < 1); with diag_bias near 1 the diagonal of
    I + S is as small as its off-diagonal part, so the dense inversions have to exchange rows."""
    rng = np.random.default_rng(500 + seed)
    m = _model(hv)
    Q, _ = np.linalg.qr(rng.normal(size=(n_block, n_block)) + 1j * rng.normal(size=(n_block, n_block)))
    tau = rng.uniform(0.05e-9, 0.4e-9, n_block)
    gain = rng.uniform(0.2, 0.7, n_block)

    def S_of(i, j):
        return lambda f: scale * np.einsum("k,kf,k->f", Q[i, :], gain[:, None] * np.exp(-2j * np.pi * np.asarray(f)[None, :] * tau[:, None]), Q[j, :]) - (diag_bias if i == j else 0.0)

    inner = [m.node() for _ in range(n_block)]
    for k in range(n_block):
        rows = chain_rows[k]
        nodes = [m.node() for _ in range(rows)] + [inner[k]]          # far end ... block node
        if port_on[k] is not None:
            m.new_port(float(rng.choice([25.0, 50.0, 75.0])), nodes[port_on[k]])
        for t in range(rows):
            kind = (k + t) % 3
            if kind == 0:
                m.inductor(nodes[t], nodes[t + 1], float(10 ** rng.uniform(-9.3, -8.5)))
            elif kind == 1:
                m.capacitor(nodes[t], nodes[t + 1], float(10 ** rng.uniform(-12.3, -11.5)))
            else:
                m.transmissionline(nodes[t], nodes[t + 1], float(rng.uniform(30, 90)), float(rng.uniform(2, 4)), float(rng.uniform(1e-3, 9e-3)))
            m.capacitor(nodes[t], m.gnd, float(10 ** rng.uniform(-12.8, -12)))
        if k % 2:
            m.resistor(inner[k], m.gnd, float(rng.uniform(200, 2000)))
    m.n_port_S(m.gnd, inner, [[S_of(i, j) for j in range(n_block)] for i in range(n_block)], 50.0)
    return m, np.linspace(0.8e9, 6e9, nF)


def build_lc_chain(hv, K, taps, coupling=None, lossy=False, seed=0):
    """Ladder of series L / shunt C (alternating kinds with `seed`), ports at `taps`; `coupling`: the two end ports sit
    behind series capacitors of that size (weakly coupled, high-Q resonator chain)."""
    rng = np.random.default_rng(seed)
    m = _model(hv)
    chain = [m.node() for _ in range(K + 1)]
    if coupling is None:
        for t in taps:
            m.new_port(50, chain[t])
    else:
        pa, pb = m.node(), m.node()
        m.new_port(50, pa)
        m.new_port(50, pb)
        m.capacitor(pa, chain[0], coupling)
        m.capacitor(chain[K], pb, coupling)
    for k in range(K):
        if coupling is not None or k % 2 == 0:
            m.inductor(chain[k], chain[k + 1], float(10 ** rng.uniform(-9.3, -8.3)) if coupling is None else 2e-9)
        else:
            m.capacitor(chain[k], chain[k + 1], float(10 ** rng.uniform(-12.5, -11.5)))
    for k in range(K + 1):
        if coupling is not None or k % 2:
            m.capacitor(chain[k], m.gnd, float(10 ** rng.uniform(-13, -11.7)) if coupling is None else 1e-12)
        else:
            m.inductor(chain[k], m.gnd, float(10 ** rng.uniform(-9, -8)))
        if lossy and k % 4 == 0:
            m.resistor(chain[k], m.gnd, float(rng.uniform(500, 5000)))
    return m
