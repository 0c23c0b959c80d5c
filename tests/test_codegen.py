"""CPU: the kernel generator behind kernel class 3 (heavi_b200/csrc/hvb_gen.cpp).

The source the library would hand to NVRTC is fetched through the host-only C-ABI entry
`hvb_codegen_source`, compiled for the CPU with g++ and a CUDA shim (tests/host_sim, test
infrastructure) and run against the oracle (oracle/mna_oracle.py, the restatement of
network.py:354-421 + solvers/spfn.py:106-179) and the reference-generated goldens.  `hvb_codegen_compile`
checks that the same text compiles for sm_100a with NVRTC (no GPU needed)."""
import os

import numpy as np
import pytest

import circuits
import heavi_b200 as hv
import host_sim
from heavi_b200 import engine
from heavi_b200 import plan as pl
from helpers import golden, parity_report, registry
from oracle import mna_oracle as orc

REG = registry(hv)


def chain_plan(m):
    return pl.compile_plan(m.stamp_table(), "auto", order="rcm")


def sim(m, f, mode=0, drivers=None, values=None):
    p = chain_plan(m)
    consts, prefs, tables, samples = pl.resolve_run(p, np.asarray(f, dtype=np.float64), drivers, values)
    gen = engine.codegen_source(p, prefs, mode, consts=consts)
    assert gen["supported"], gen["reason"]
    S, f32, status = host_sim.run(gen, p.P, f, consts, prefs, tables, samples, p)
    return S, f32, status, gen


def oracle_S(m, f):
    return orc.run_sp(m.records(), m.n_nodes, len(m.sources), np.asarray(f, dtype=np.float64), method="lu_batched")


CHAINS = ["c1_readme", "c2_bandpass", "c2_lowpass", "c2_highpass", "ka_divider", "ka_line"]


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("name", CHAINS)
def test_generated_kernel_matches_reference_golden(name, mode):
    g = golden(name)
    m = REG[name]()
    f = g["f"][:: max(1, len(g["f"]) // 400)]
    sel = np.arange(len(g["f"]))[:: max(1, len(g["f"]) // 400)]
    S, f32, status, gen = sim(m, f, mode)
    assert status == 0 and gen["rolled"] == (mode == 2)
    assert np.array_equal(f32, g["f32"][sel])
    frac, adjudicated, ext = parity_report(S[0], g["S"][..., sel], g["S_ext"][..., sel])
    assert frac >= 0.99 and adjudicated and ext < 5e-14, (frac, adjudicated, ext)


def test_rolled_and_straight_line_forms_agree_bit_for_bit():
    m, f = circuits.build_c5(hv, nF=24)
    S1, _, st1, g1 = sim(m, f, 1)
    S2, _, st2, g2 = sim(m, f, 2)
    assert st1 == 0 and st2 == 0 and not g1["rolled"] and g2["rolled"]
    assert g2["n_cases"] < 60 and len(g2["rows"]) == 4 * 190
    assert np.array_equal(S1, S2)
    g = golden("c5_ladder")
    S, _, st, _ = sim(m, g["f"], 0)
    frac, adjudicated, ext = parity_report(S[0], g["S"], g["S_ext"])
    assert frac == 1.0 and ext < 5e-14


@pytest.mark.parametrize("n_sections,taps", [(3, [0]), (3, [1]), (3, [2]), (4, [0, 3]), (5, [0, 1, 2, 3, 4]), (9, [2, 3, 7]),
                                             (12, [0, 5, 6, 11]), (30, [1, 2, 3, 4, 5, 20, 28, 29]),
                                             (44, list(range(0, 44, 4))), (64, [63]), (33, list(range(16)))])
def test_port_positions_and_counts(n_sections, taps):
    """Ports anywhere along the chain, adjacent ports, one port, 16 ports; both code forms."""
    rng = np.random.default_rng(n_sections * 7 + len(taps))
    m = hv.Model(suppress_loadbar=True)
    nodes = [m.node() for _ in range(n_sections)]
    for t in taps:
        m.new_port(float(rng.choice([25.0, 50.0, 75.0])), nodes[t])
    for k in range(n_sections - 1):
        kind = k % 3
        if kind == 0:
            m.inductor(nodes[k], nodes[k + 1], float(10 ** rng.uniform(-9.3, -8.3)))
        elif kind == 1:
            m.transmissionline(nodes[k], nodes[k + 1], float(rng.uniform(30, 90)), float(rng.uniform(2, 4)), float(rng.uniform(1e-3, 9e-3)))
        else:
            m.resistor(nodes[k], nodes[k + 1], float(rng.uniform(5, 80)))
    for k in range(n_sections):
        m.capacitor(nodes[k], m.gnd, float(10 ** rng.uniform(-12.5, -11.5)))
    f = np.linspace(0.4e9, 7e9, 23)
    ref = oracle_S(m, f)
    for mode in (1, 2, 3):
        S, _, status, gen = sim(m, f, mode)
        assert status == 0
        assert np.max(np.abs(S[0] - ref)) < 2e-13, (mode, gen["two_sided"], np.max(np.abs(S[0] - ref)))


def _lc_chain(K, taps, coupling=None, lossy=False, seed=0):
    return circuits.build_lc_chain(hv, K, taps, coupling, lossy, seed)


def test_two_sided_form_on_config5_matches_the_reference_golden_with_less_work():
    """Long chains are walked from both ends towards a cut (hvb_gen.cpp header): same S as the one-sided row loop within
    rounding, checked against the reference's own output; the border work drops by a quarter."""
    m, _ = circuits.build_c5(hv, nF=8)
    g = golden("c5_ladder")
    S3, _, st3, g3 = sim(m, g["f"], 3)
    S2, _, st2, g2 = sim(m, g["f"], 2)
    S0, _, st0, g0 = sim(m, g["f"], 0)
    assert st3 == 0 and st2 == 0 and st0 == 0
    assert g3["two_sided"] and g0["two_sided"] and not g2["two_sided"] and g3["cut_row"] == 95 and g2["cut_row"] == -1
    assert np.array_equal(S0, S3)                                  # auto picks the two-sided form for this chain
    assert g3["flops_per_solve"] < 0.8 * g2["flops_per_solve"]
    assert len(g3["rows"]) == 4 * 190 and g3["n_cases"] < 70
    frac, adjudicated, ext = parity_report(S3[0], g["S"], g["S_ext"])
    assert frac == 1.0 and ext < 5e-14
    assert np.max(np.abs(S3 - S2)) < 5e-14


@pytest.mark.parametrize("K,taps,lossy", [(12, [0, 12], False), (40, [0, 40], False), (60, [0, 30, 60], False), (60, [0, 20, 40, 60], True),
                                          (45, [3, 4, 40, 41, 42], True), (50, [0, 1, 2, 3, 4, 5, 6, 50], True)])
def test_two_sided_form_on_lossless_and_lossy_ladders(K, taps, lossy):
    """The cut grounds each half at the other's first node; lossless LC ladders (resonant halves) are the hard case."""
    m = _lc_chain(K, taps, lossy=lossy, seed=K)
    f = np.linspace(0.2e9, 8e9, 257)
    ref = oracle_S(m, f)
    S3, _, st3, g3 = sim(m, f, 3)
    S2, _, st2, _ = sim(m, f, 2)
    assert g3["two_sided"] and st2 == 0
    tol = 1e-12 + 1e-10 * np.abs(ref)
    assert np.all(np.abs(S2[0] - ref) <= tol)
    if st3 == 0:
        assert np.all(np.abs(S3[0] - ref) <= tol), np.max(np.abs(S3[0] - ref))
    else:
        assert st3 == engine.HVB_STATUS_RERUN                      # refused, never wrong silently


@pytest.mark.parametrize("seed", range(8))
def test_two_sided_form_fuzz_against_oracle(seed):
    """Random chains (41-110 rows, 2-8 ports anywhere, random mix of inductors, capacitors, resistors, lossless and lossy
    lines, some shunt loss) on the two-sided row loop: within the parity gate of the oracle, or refused (never wrong)."""
    rng = np.random.default_rng(9000 + seed)
    K = int(rng.integers(41, 111))
    m = hv.Model(suppress_loadbar=True)
    chain = [m.node() for _ in range(K + 1)]
    taps = sorted(set(int(t) for t in rng.integers(0, K + 1, int(rng.integers(2, 9)))))
    for t in taps:
        m.new_port(float(rng.choice([25.0, 50.0, 75.0])), chain[t])
    for k in range(K):
        kind = int(rng.integers(0, 5))
        if kind == 0:
            m.inductor(chain[k], chain[k + 1], float(10 ** rng.uniform(-9.5, -8.3)))
        elif kind == 1:
            m.capacitor(chain[k], chain[k + 1], float(10 ** rng.uniform(-12.3, -11)))
        elif kind == 2:
            m.resistor(chain[k], chain[k + 1], float(rng.uniform(1, 60)))
        elif kind == 3:
            m.transmissionline(chain[k], chain[k + 1], float(rng.uniform(30, 90)), float(rng.uniform(2, 4)), float(rng.uniform(1e-3, 9e-3)))
        else:
            m.transmissionline(chain[k], chain[k + 1], float(rng.uniform(30, 90)), complex(rng.uniform(2, 4), -rng.uniform(0.01, 0.1)), float(rng.uniform(1e-3, 9e-3)))
    for k in range(K + 1):
        m.capacitor(chain[k], m.gnd, float(10 ** rng.uniform(-12.8, -11.8)))
        if rng.random() < 0.2:
            m.resistor(chain[k], m.gnd, float(rng.uniform(300, 5000)))
    f = np.sort(10 ** rng.uniform(8.5, 9.9, 24))
    ref = oracle_S(m, f)
    S3, _, st3, g3 = sim(m, f, 3)
    assert g3["rolled"]
    if not g3["two_sided"]:                                       # no admissible cut for this draw: the one-sided loop serves it
        assert len(taps) < 2 or st3 == 0
    if st3 == 0:
        assert np.all(np.abs(S3[0] - ref) <= 1e-12 + 1e-10 * np.abs(ref)), np.max(np.abs(S3[0] - ref))
    else:
        assert g3["two_sided"] and st3 == engine.HVB_STATUS_RERUN
        S2, _, st2, _ = sim(m, f, 2)
        assert st2 == 0 and np.all(np.abs(S2[0] - ref) <= 1e-12 + 1e-10 * np.abs(ref))


def test_two_sided_form_needs_a_cut_between_ports_and_auto_mode_needs_a_saving():
    f = np.linspace(1e9, 3e9, 5)
    # one port, or all ports on one side of every admissible cut: no cut -> mode 3 quietly gives the one-sided row loop
    for taps in ([7], [0, 1], [48, 49, 50]):
        m = _lc_chain(50, taps, lossy=True)
        S, _, st, gen = sim(m, f, 3)
        assert gen["rolled"] and not gen["two_sided"] and st == 0
        assert np.max(np.abs(S[0] - oracle_S(m, f))) < 2e-13
    # every row a port row: no cut either
    m = _lc_chain(15, list(range(16)), lossy=True)
    assert not sim(m, f, 3)[3]["two_sided"]
    # a port at the far end only: cutting saves nothing (its recurrence is one row long already) -> auto stays one-sided
    m = _lc_chain(50, [49, 50], lossy=True)
    assert not sim(m, f, 0)[3]["two_sided"]
    m = _lc_chain(50, [0, 50], lossy=True)                       # ports at the two ends: 52 tap-rows either way
    assert not sim(m, f, 0)[3]["two_sided"]
    m = _lc_chain(50, [0, 10, 50], lossy=True)                   # 93 tap-rows one-sided, 55 with the cut next to the middle port
    gen = sim(m, f, 0)[3]
    assert gen["two_sided"] and gen["cut_row"] in (11, 38)


def test_two_sided_form_refuses_an_unsafe_cut():
    """Weakly coupled resonator chain: at the halves' resonances the 2 x 2 interface solve cancels beyond 10^3 and the kernel
    sets HVB_STATUS_RERUN (the library then repeats the call one-sided, tests/test_gpu_generated.py)."""
    m = _lc_chain(44, None, coupling=1e-15)
    f = np.linspace(0.2e9, 8e9, 4000)
    S3, _, st3, g3 = sim(m, f, 3)
    S2, _, st2, g2 = sim(m, f, 2)
    assert g3["two_sided"] and st3 == engine.HVB_STATUS_RERUN and st2 == 0
    ref = oracle_S(m, f)
    assert np.all(np.abs(S2[0] - ref) <= 1e-12 + 1e-10 * np.abs(ref))


def test_monte_carlo_samples_tables_and_lossy_lines():
    m, f, mc = circuits.build_c4(hv, nF=31)
    np.random.seed(3)
    values = mc.draw(5)
    S, _, status, gen = sim(m, f, 0, mc._random_numbers, values)
    assert status == 0 and S.shape == (5, 2, 2, 31)
    for s in range(5):
        for r, v in zip(mc._random_numbers, values[s]):
            r._value = float(v)
        assert np.max(np.abs(S[s] - oracle_S(m, f))) < 1e-13
    # host-tabulated callable, complex permittivity (lossy line), sample-valued port impedance stays out
    m2 = hv.Model(suppress_loadbar=True)
    a, b = m2.quick_port(50), m2.quick_port(75)
    mid, mid2 = m2.node(), m2.node()
    m2.impedance(a, mid, lambda fr: 3.0 + 1j * 2e-9 * fr)
    m2.transmissionline(mid, mid2, 60.0, 3.1 - 0.05j, 0.011)
    m2.admittance(mid2, b, lambda fr: 0.02 + 1j * 1e-12 * fr)
    m2.capacitor(mid, m2.gnd, 0.4e-12)
    f2 = np.linspace(1e9, 5e9, 17)
    S2, _, status, _ = sim(m2, f2, 0)
    assert status == 0 and np.max(np.abs(S2[0] - oracle_S(m2, f2))) < 1e-13


@pytest.mark.parametrize("mode", [2, 3])
def test_long_chain_with_samples_tables_and_lossy_lines_in_the_row_loop(mode):
    """The row loop (one- and two-sided) with every parameter kind on a 60-section chain: Monte-Carlo sample columns,
    host-tabulated callables, lossy lines (complex permittivity: the general, non-inlined line form) next to lossless
    ones (the closed form), ports in the middle of the chain."""
    rng = np.random.default_rng(77)
    m = hv.Model(suppress_loadbar=True)
    mc = hv.MonteCarlo()
    K = 60
    chain = [m.node() for _ in range(K + 1)]
    for t in (0, 17, 18, 41, 60):
        m.new_port(50.0, chain[t])
    for k in range(K):
        kind = k % 4
        if kind == 0:
            m.inductor(chain[k], chain[k + 1], mc.gaussian(float(10 ** rng.uniform(-9.3, -8.5)), 2e-11))
        elif kind == 1:
            m.transmissionline(chain[k], chain[k + 1], float(rng.uniform(30, 90)), float(rng.uniform(2, 4)), float(rng.uniform(1e-3, 9e-3)))
        elif kind == 2:
            m.transmissionline(chain[k], chain[k + 1], float(rng.uniform(30, 90)), complex(rng.uniform(2, 4), -0.04), float(rng.uniform(1e-3, 9e-3)))
        else:
            m.impedance(chain[k], chain[k + 1], lambda fr, r=float(rng.uniform(2, 20)): r + 1j * 1.5e-9 * fr)
    for k in range(K + 1):
        m.capacitor(chain[k], m.gnd, mc.uniform(0.2e-12, 0.6e-12) if k % 7 == 0 else float(10 ** rng.uniform(-12.6, -12)))
    f = np.linspace(0.7e9, 5.5e9, 19)
    np.random.seed(5)
    values = mc.draw(3)
    S, _, status, gen = sim(m, f, mode, mc._random_numbers, values)
    assert status == 0 and gen["rolled"] and gen["two_sided"] == (mode == 3) and S.shape == (3, 5, 5, 19)
    assert "gen_tline_lossless" in gen["src"] and "gen_tline_chain(" in gen["src"]
    for s_ in range(3):
        for r, v in zip(mc._random_numbers, values[s_]):
            r._value = float(v)
        ref = oracle_S(m, f)
        assert np.all(np.abs(S[s_] - ref) <= 1e-12 + 1e-10 * np.abs(ref)), (s_, np.max(np.abs(S[s_] - ref)))


def test_singular_and_nonfinite_points_set_status_bits():
    m = hv.Model(suppress_loadbar=True)
    a, b = m.quick_port(50), m.quick_port(50)
    mid = m.node()
    m.inductor(a, mid, 1e-9)
    m.inductor(mid, b, 1e-9)
    _, _, status, _ = sim(m, np.array([0.0, 1e9]))
    assert status & engine.HVB_STATUS_NONFINITE


def test_unsupported_structures_are_refused_with_a_reason():
    m, f = circuits.build_c3(hv, nF=8)                          # 8-port block: not a chain, but a hub (kernel E)
    p = pl.compile_plan(m.stamp_table(), "auto")
    gen = engine.codegen_source(p, pl.resolve_run(p, f)[1])
    assert gen["supported"] and gen["hub"]
    mb, _ = circuits.build_c2(hv, "bandstop", 8)                # shunt series-LC branches: a tree, neither chain nor hub
    genb = engine.codegen_source(pl.compile_plan(mb.stamp_table(), "auto", order="natural"))
    assert not genb["supported"] and genb["reason"]
    m2, _ = circuits.build_band_ladder(hv, K=40, reach=2, nports=2)  # pentadiagonal
    gen2 = engine.codegen_source(pl.compile_plan(m2.stamp_table(), "auto", order="rcm"))
    assert not gen2["supported"] and "tridiagonal" in gen2["reason"]
    m3, _ = circuits.build_c2(hv, "bandpass", 8)                 # full MNA keeps the source rows: general epilogue
    gen3 = engine.codegen_source(pl.compile_plan(m3.stamp_table(), "full", order="rcm"))
    assert not gen3["supported"]


@pytest.mark.parametrize("name,mode", [("c2_bandpass", 1), ("c5_ladder", 2), ("c5_ladder", 3)])
def test_generated_source_compiles_for_sm100a(name, mode):
    m = REG[name]()
    p = chain_plan(m)
    gen = engine.codegen_source(p, None, mode)
    old = os.environ.get("HVB_NVRTC_FLAGS")
    os.environ["HVB_NVRTC_FLAGS"] = "--ptxas-options=-v"
    try:
        log = engine.codegen_compile(gen["src"])
    finally:
        if old is None:
            os.environ.pop("HVB_NVRTC_FLAGS")
        else:
            os.environ["HVB_NVRTC_FLAGS"] = old
    assert "hvb_gen_kernel" in log and "registers" in log


# ---------------------------------------------------------------------------------------------------
# kernel E: hub netlists (one N-port S block + chains), heavi_b200/csrc/hvb_gen_hub.cpp
# ---------------------------------------------------------------------------------------------------
def hub_sim(m, f, drivers=None, values=None):
    p = pl.compile_plan(m.stamp_table(), "auto")
    consts, prefs, tables, samples = pl.resolve_run(p, np.asarray(f, dtype=np.float64), drivers, values)
    gen = engine.codegen_source(p, prefs, 0, consts=consts)
    assert gen["supported"] and gen["hub"], gen["reason"]
    S, f32, status = host_sim.run(gen, p.P, f, consts, prefs, tables, samples, p)
    return S, f32, status, gen


def test_hub_kernel_config3_matches_reference_golden():
    g = golden("c3_touchstone")
    m = REG["c3_touchstone"]()
    S, f32, status, gen = hub_sim(m, g["f"])
    assert status == 0 and gen["smem_bytes"] == 8 * 8 * 32 * 16 and gen["threads"] == 32      # the 8 x 8 matrix alone: 1 KB per thread
    assert np.array_equal(f32, g["f32"])
    frac, adjudicated, ext = parity_report(S[0], g["S"], g["S_ext"])
    assert frac == 1.0 and adjudicated and ext < 5e-14


@pytest.mark.parametrize("n_block,chain_rows,port_on", [
    (4, (2, 1, 0, 3), (0, 0, None, 1)),            # chains of every length, a port in the middle of a chain, a bare node
    (2, (1, 1), (0, 0)),
    (3, (0, 2, 0), (0, 1, 0)),                     # ports directly on block nodes
    (8, (1,) * 8, (0,) * 8),
    (5, (4, 0, 2, 0, 1), (2, None, 0, 0, None)),   # stubs without ports
])
def test_hub_kernel_structures_against_oracle(n_block, chain_rows, port_on):
    m, f = circuits.build_hub(hv, n_block, chain_rows, port_on, seed=n_block)
    S, _, status, gen = hub_sim(m, f)
    assert status == 0
    assert np.max(np.abs(S[0] - oracle_S(m, f))) < 2e-13


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_hub_kernel_inversions_exchange_rows_when_they_must(seed):
    """S -> 0.3 S - 0.995 I: the diagonal of I + S is no larger than its off-diagonal part (90-140 row exchanges over the 41 points
    for these seeds, counted with the same rule on the host), so the in-place Gauss-Jordan
    inversions (hvb_gen_hub.cpp, gen_dense_inv) have to pivot, and the column un-scrambling afterwards matters."""
    m, f = circuits.build_hub(hv, 6, (1, 2, 0, 1, 0, 2), (0, 1, 0, 0, None, 0), seed=seed, nF=41, scale=0.3, diag_bias=0.995)
    S, _, status, gen = hub_sim(m, f)
    ref = oracle_S(m, f)
    assert status == 0
    assert np.all(np.abs(S[0] - ref) <= 1e-12 + 1e-10 * np.abs(ref)), np.max(np.abs(S[0] - ref))


@pytest.mark.parametrize("seed", range(6))
def test_hub_kernel_fuzz_against_oracle(seed):
    """Random hubs: 2-8 block ports, chains of 0-4 rows, ports anywhere or nowhere on them, half of the draws with a block
    whose I + S needs row exchanges."""
    rng = np.random.default_rng(77000 + seed)
    nb = int(rng.integers(2, 9))
    rows = tuple(int(x) for x in rng.integers(0, 5, nb))
    port_on = [None if rng.random() < 0.25 else (int(rng.integers(0, rows[k] + 1)) if rows[k] > 0 else 0) for k in range(nb)]
    if all(p is None for p in port_on):
        port_on[0] = 0
    sc, db = (1.0, 0.0) if rng.random() < 0.5 else (float(rng.uniform(0.05, 0.5)), float(rng.uniform(0.8, 1.0)))
    m, f = circuits.build_hub(hv, nb, rows, tuple(port_on), seed=seed, nF=9, scale=sc, diag_bias=db)
    S, _, status, gen = hub_sim(m, f)
    ref = oracle_S(m, f)
    assert status in (0, engine.HVB_STATUS_ZERO_PIVOT)            # refused (unsafe boundary pivot) is legal, wrong is not
    if status == 0:
        assert np.all(np.abs(S[0] - ref) <= 1e-12 + 1e-10 * np.abs(ref)), np.max(np.abs(S[0] - ref))


def test_hub_kernel_with_monte_carlo_samples_and_tabulated_parts():
    """Kernel E with sample columns and host-tabulated callables on the chains, a port on a block node, a lossy line."""
    rng = np.random.default_rng(9)
    m = hv.Model(suppress_loadbar=True)
    mc = hv.MonteCarlo()
    nb = 4
    Q, _ = np.linalg.qr(rng.normal(size=(nb, nb)) + 1j * rng.normal(size=(nb, nb)))
    tau = rng.uniform(0.05e-9, 0.4e-9, nb)
    gain = rng.uniform(0.2, 0.7, nb)

    def S_of(i, j):
        return lambda fr: np.einsum("k,kf,k->f", Q[i, :], gain[:, None] * np.exp(-2j * np.pi * np.asarray(fr)[None, :] * tau[:, None]), Q[j, :])

    inner = [m.node() for _ in range(nb)]
    a, a2 = m.node(), m.node()
    m.new_port(50.0, a)
    m.inductor(a, a2, mc.gaussian(1.2e-9, 4e-11))
    m.capacitor(a2, m.gnd, mc.uniform(0.3e-12, 0.7e-12))
    m.capacitor(a2, inner[0], 2.1e-12)
    b = m.node()
    m.new_port(75.0, b)
    m.transmissionline(b, inner[1], 55.0, 2.9 - 0.03j, 0.007)
    m.new_port(50.0, inner[2])                                   # a port directly on a block node
    c = m.node()
    m.impedance(c, inner[3], lambda fr: 4.0 + 1j * 1e-9 * fr)    # a stub without a port
    m.capacitor(c, m.gnd, mc.gaussian(0.8e-12, 2e-14))
    m.n_port_S(m.gnd, inner, [[S_of(i, j) for j in range(nb)] for i in range(nb)], 50.0)
    f = np.linspace(0.9e9, 5e9, 21)
    np.random.seed(12)
    values = mc.draw(4)
    S, _, status, gen = hub_sim(m, f, mc._random_numbers, values)
    assert status == 0 and gen["hub"] and S.shape == (4, 3, 3, 21)
    for s_ in range(4):
        for r, v in zip(mc._random_numbers, values[s_]):
            r._value = float(v)
        ref = oracle_S(m, f)
        assert np.all(np.abs(S[s_] - ref) <= 1e-12 + 1e-10 * np.abs(ref)), (s_, np.max(np.abs(S[s_] - ref)))


def test_hub_kernel_flags_an_unsafe_boundary_pivot_and_refuses_other_structures():
    # a chain whose last row is an (almost) ideal series resonator next to the block node: the restricted pivot is
    # unsafe there, which must raise the zero-pivot bit (the host then re-runs on the dense kernel)
    m = hv.Model(suppress_loadbar=True)
    b, c, mid = m.node(), m.node(), m.node()
    m.new_port(50.0, b)
    m.new_port(50.0, c)
    C2, f0 = 2e-12, 1e9
    m.capacitor(mid, b, C2)                           # stub on block node b: series C into a node ...
    m.inductor(mid, m.gnd, 1.0 / ((2 * np.pi * f0) ** 2 * C2))   # ... whose only other part resonates with it at f0
    m.n_port_S(m.gnd, [b, c], [[lambda f: 0.1 + 0 * f, lambda f: 0.7 + 0 * f], [lambda f: 0.7 + 0 * f, lambda f: 0.2 + 0 * f]], 50.0)
    S, _, status, _ = hub_sim(m, np.array([f0, 1.7e9]))
    assert status & engine.HVB_STATUS_ZERO_PIVOT
    S2, _, status2, _ = hub_sim(m, np.array([1.3e9, 1.7e9]))
    assert status2 == 0 and np.max(np.abs(S2[0] - oracle_S(m, np.array([1.3e9, 1.7e9])))) < 1e-13
    # two blocks, or a mesh around the block: not a hub
    m2, _ = circuits.build_hub(hv, 3, (1, 1, 1), (0, 0, 0))
    nodes = [n for n in m2._nodes][-2:]
    m2.resistor(m2._nodes[4], m2._nodes[6], 30.0)     # ties two chains together
    gen = engine.codegen_source(pl.compile_plan(m2.stamp_table(), "auto"))
    assert not gen["supported"] and "hub kernel" in gen["reason"]
