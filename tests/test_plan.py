"""CPU: the plan compiler (value program + cell program + port tables) interpreted in numpy must
reproduce the reference's dense tensor and S for every golden circuit, in every solver mode."""
import numpy as np
import pytest

import heavi_b200 as hv
import plan_interp as pi
from heavi_b200 import plan as pl
from helpers import dense_y, golden, max_ulps, parity_report, registry

REG = registry(hv)


@pytest.mark.parametrize("name", sorted(REG))
def test_dump_plan_reproduces_reference_tensor(name):
    g = golden(name)
    tab = REG[name]().stamp_table()
    W = pi.dump_plan(pl.compile_plan(tab, "dump"), g["y_f"])
    Yg = dense_y(g)
    assert np.array_equal(W != 0, Yg != 0) or max_ulps(W, Yg) < 200    # same stamp pattern
    # inversion-based N-port blocks and lossy lines differ by rounding only
    assert np.max(np.abs(W - Yg)) <= 1e-13 * max(1.0, np.max(np.abs(Yg)))


@pytest.mark.parametrize("mode", ["full", "norton"])
@pytest.mark.parametrize("name", sorted(REG))
def test_plan_solves_to_reference(name, mode):
    g = golden(name)
    tab = REG[name]().stamp_table()
    if mode == "norton" and not pl.norton_legal(tab):
        pytest.skip("port internal node is shared")
    p = pl.compile_plan(tab, mode)
    f = g["f"] if name != "c5_ladder" else g["f"][:4]
    S = pi.run_plan(p, f)[0]
    nf = len(f)
    frac, adjudicated, ext = parity_report(S, g["S"][:, :, :nf], g["S_ext"][:, :, :nf])
    assert frac >= 0.998 and adjudicated and ext < 5e-14
    expected_n = tab.N + tab.M - 1 if mode == "full" else tab.N - 1 - tab.P
    assert p.n == expected_n and p.ncols == p.n + tab.P


def test_norton_illegal_when_internal_node_is_shared():
    m = hv.Model(suppress_loadbar=True)
    a = m.node()
    p1 = m.new_port(50, a)
    b = m.quick_port(50)
    m.resistor(a, b, 10.0)
    m.capacitor(p1.int_node, m.gnd, 1e-12)       # something else touches the port's internal node
    tab = m.stamp_table()
    assert not pl.norton_legal(tab)
    assert pl.compile_plan(tab, "auto").mode == "full"
    with pytest.raises(ValueError):
        pl.compile_plan(tab, "norton")


def test_duplicate_index_last_write_wins():
    """App. A: a part with both terminals on one node adds +y (not 0) to the diagonal."""
    m = hv.Model(suppress_loadbar=True)
    a = m.quick_port(50)
    m.resistor(a, a, 25.0)
    m.capacitor(a, m.gnd, 1e-12)
    p = pl.compile_plan(m.stamp_table(), "dump")
    W = pi.dump_plan(p, np.array([1e9]))
    ia = a.index
    assert W[ia, ia, 0].real == pytest.approx(1 / 50 + 1 / 25)


def test_callable_scalar_and_table_parameters():
    m = hv.Model(suppress_loadbar=True)
    a = m.quick_port(50)
    b = m.quick_port(50)
    m.impedance(a, b, lambda f: 30.0)                         # callable returning a scalar
    m.admittance(b, m.gnd, lambda f: 1e-3 + 2j * np.pi * f * 1e-12)
    p = pl.compile_plan(m.stamp_table(), "auto")
    f = np.array([1e9, 2e9, 3e9])
    consts, prefs, tables, samples = pl.resolve_run(p, f)
    assert tables.shape == (1, 3)                             # only the genuinely f-dependent one
    kinds = [k for k, _ in prefs]
    assert kinds.count(pl.PK_TABLE) == 1
    S = pi.run_plan(p, f)[0]
    from oracle import mna_oracle as orc
    ref = orc.run_sp(m.records(), m.n_nodes, len(m.sources), f, method="lu_batched")
    assert np.max(np.abs(S - ref)) < 1e-13


def test_sample_parameters_and_batch_columns():
    import circuits
    g = golden("c4_mc")
    m, f, mc = circuits.build_c4(hv)
    p = pl.compile_plan(m.stamp_table(), "auto")
    assert p.n_samples_cols == 10
    S = pi.run_plan(p, g["f"][:64], mc._random_numbers, g["mc_draws"])
    assert S.shape[0] == 4
    frac, adj, ext = parity_report(S, g["mc_S"][:, :, :, :64], g["mc_S"][:, :, :, :64])
    assert frac >= 0.95
    assert np.max(np.abs(S - g["mc_S"][:, :, :, :64])) < 1e-10


@pytest.mark.parametrize("name", ["c1_readme", "c2_lowpass", "c3_touchstone", "rnd_05", "rnd_19"])
def test_padded_plan_is_equivalent(name):
    """Identity-row / zero-column padding (what the register kernel's fixed shapes need) must not
    change the solution, and the per-row entry lists must describe the same matrix as the cell CSR."""
    g = golden(name)
    tab = REG[name]().stamp_table()
    base = pl.compile_plan(tab, "auto")
    padded = pl.compile_plan(tab, base.mode, pad_to=(base.n + 3, base.P + 2))
    assert padded.n_real == base.n and padded.flops_per_solve() == base.flops_per_solve()
    f = g["f"][:32]
    assert np.array_equal(pi.run_plan(padded, f), pi.run_plan(base, f)) or \
        np.max(np.abs(pi.run_plan(padded, f) - pi.run_plan(base, f))) < 1e-14
    # row lists == cell CSR
    for p in (base, padded):
        cells = {}
        for r in range(p.n):
            col_seen = []
            for e in p.row_ent[p.row_ptr[r]:p.row_ptr[r + 1]]:
                e = int(e) & 0xFFFFFFFF
                c, vid, neg, last = (e >> 20) & 0x7FF, (e >> 1) & 0x7FFFF, e & 1, e >> 31
                cells.setdefault((r, c), []).append((vid << 1) | neg)
                col_seen.append((c, last))
            assert [c for c, _ in col_seen] == sorted(c for c, _ in col_seen)
        for r in range(p.n):
            for c in range(p.ncols):
                lo, hi = p.cell_ptr[r * p.ncols + c], p.cell_ptr[r * p.ncols + c + 1]
                assert list(p.cell_ent[lo:hi]) == cells.get((r, c), [])


@pytest.mark.parametrize("name", ["c1_readme", "c2_bandpass", "c2_highpass", "c3_touchstone", "c5_ladder", "rnd_04"])
def test_closed_form_epilogue_equals_reference_formula(name):
    """Norton plan + ground-referenced constant-Z0 ports: the closed form must equal spfn.py:159-176."""
    g = golden(name)
    tab = REG[name]().stamp_table()
    p = pl.compile_plan(tab, "norton")
    assert p.epi_simple == 1 and sorted(x for x in p.epi_port_of_row if x >= 0) == list(range(p.P))
    f = g["f"][:8]
    consts, prefs, tables, samples = pl.resolve_run(p, f)
    a = pi.Interp(p, f, consts, prefs, tables, samples[0])
    general = a.solve()
    a.use_closed_form = True
    assert np.max(np.abs(a.solve() - general)) < 4e-15


def test_closed_form_epilogue_not_used_when_port_is_not_ground_referenced():
    import circuits
    for seed in range(48):
        m, _ = circuits.build_random(hv, seed)
        tab = m.stamp_table()
        if not pl.norton_legal(tab):
            continue
        p = pl.compile_plan(tab, "norton")
        grounded = all(int(r[3]) == 0 for r in tab.port_table)
        distinct = len({int(r[1]) for r in tab.port_table}) == tab.P
        assert p.epi_simple == (1 if grounded and distinct else 0)
    assert pl.compile_plan(REG["c1_readme"]().stamp_table(), "full").epi_simple == 0


def _half_bandwidth(p):
    return max((abs(r - c) for r in range(p.n) for c in range(p.n)
                if p.cell_ptr[r * p.ncols + c + 1] > p.cell_ptr[r * p.ncols + c]), default=0)


@pytest.mark.parametrize("reach,K,nports,mode", [(1, 40, 2, "auto"), (2, 60, 3, "auto"), (4, 80, 3, "auto"),
                                                  (2, 50, 3, "full"), (1, 33, 2, "full")])
def test_band_ordering_is_a_symmetric_permutation(reach, K, nports, mode):
    """The bandwidth-reducing row order (what the banded kernel needs) renumbers the unknowns and
    changes nothing else: same S as the natural order, structure inside the reported band."""
    import circuits
    m, f = circuits.build_band_ladder(hv, K=K, reach=reach, nports=nports, seed=reach * 100 + K)
    tab = m.stamp_table()
    nat = pl.compile_plan(tab, mode, order="natural")
    rcm = pl.compile_plan(tab, mode, order="rcm")
    assert nat.band == -1 and rcm.band >= reach and rcm.band <= pl.BAND_MAX_HALF_WIDTH
    assert _half_bandwidth(rcm) == rcm.band
    if mode == "auto":
        assert rcm.band == reach                       # the netlist's own chain order is already optimal
    assert sorted(rcm.row_ent.tolist()) != [] and rcm.n == nat.n and len(rcm.row_ent) == len(nat.row_ent)
    f = f[:6]
    assert np.max(np.abs(pi.run_plan(rcm, f) - pi.run_plan(nat, f))) < 1e-12
    auto = pl.compile_plan(tab, mode)                  # "auto" orders systems beyond the register kernels
    assert (auto.band >= 0) == (auto.n_real >= pl.BAND_MIN_ROWS)


def test_band_ordering_skips_dense_and_small_systems():
    g = golden("c3_touchstone")
    assert pl.compile_plan(REG["c3_touchstone"]().stamp_table()).band == -1      # N-port block, n = 24
    assert pl.compile_plan(REG["c2_bandpass"]().stamp_table()).band == -1        # n = 7: register kernel
    c5 = pl.compile_plan(REG["c5_ladder"]().stamp_table())
    assert c5.band == 1 and c5.n_real == 191 and _half_bandwidth(c5) == 1
    with pytest.raises(ValueError):
        pl.compile_plan(REG["c5_ladder"]().stamp_table(), order="nearest")


def _smd_model():
    """All three surface-mount models in one two-port network."""
    from heavi_b200.lib import smd
    m = hv.Model(suppress_loadbar=True)
    a, b = m.quick_port(50), m.quick_port(75)
    n1, n2 = m.node(), m.node()
    smd.SMDInductor(3.3e-9, smd.SMDInductorSize.L0402).connect(a, n1)
    smd.SMDCapacitor(1.8e-12, smd.SMDCapacitorSize.C0603).connect(n1, m.gnd)
    smd.SMDResistor(22.0, smd.SMDResistorSize.R0402).connect(n1, n2)
    smd.SMDCapacitor(4.7e-12, smd.SMDCapacitorSize.C0402).connect(n2, b)
    smd.SMDInductor(10e-9, smd.SMDInductorSize.L0603).connect(b, m.gnd)
    return m


def test_smd_closed_forms_are_compiled_not_tabulated():
    """lib/smd.py parts describe themselves as closed forms (PK_SMD): nothing is tabulated on the host, and
    the interpreter's / device's formula reproduces the part's own callable bit for bit."""
    from oracle import mna_oracle as orc
    m = _smd_model()
    tab = m.stamp_table()
    f = np.linspace(0.3e9, 7e9, 23)
    for mode in ("dump", "full", "norton"):
        p = pl.compile_plan(tab, mode)
        assert len(p.table_params) == 0 and sum(int(k) == pl.PK_SMD for k, _ in p.prefs) == 5
    W = pi.dump_plan(pl.compile_plan(tab, "dump"), f)
    Y = orc.assemble(m.records(), m.n_nodes + len(m.sources), f)
    assert np.array_equal(W, Y)                                   # same numpy arithmetic, same bits
    S = pi.run_plan(pl.compile_plan(tab), f)[0]
    So = orc.run_sp(m.records(), m.n_nodes, len(m.sources), f, method="lu_batched")
    assert np.max(np.abs(S - So)) < 1e-13
    # a finite `inf` clamp is not part of the closed form: such a parameter falls back to a table
    from heavi_b200.params import SMDImpedance
    z = SMDImpedance(2, 0.1, 0.5e-9, 1e-12, lambda fr: 0.1 + 1j * fr)
    z._inf = 1e30
    assert z.describe()[0] == "table"


def _coupler_model():
    """Every part of lib/tl.py in one four-port network."""
    from heavi_b200.lib import tl
    m = hv.Model(suppress_loadbar=True)
    p = [m.quick_port(50) for _ in range(4)]
    a, b, c, d = (m.node() for _ in range(4))
    tl.Wilkinson(50.0, 2.4e9, er=2.9).connect(p[0], a, b)
    tl.BranchlineCoupler(50.0, 2.4e9, er=3.4).connect(a, p[1], c, d)
    tl.RatraceCoupler(50.0, 2.4e9).connect(b, c, p[2], d)
    tl.Transformer(50.0, 92.0, 2.4e9, N_sections=3, ripple=0.1, er=2.2).connect(d, p[3])
    m.resistor(c, m.gnd, 180.0)
    return m


def test_coupler_library_compiles_to_closed_forms():
    """lib/tl.py parts: their phase constants are closed forms (LinePhase / PhaseConstant), nothing is
    tabulated, and the compiled stamps equal the host evaluation of the same callables."""
    from oracle import mna_oracle as orc
    m = _coupler_model()
    tab = m.stamp_table()
    f = np.linspace(1.2e9, 3.6e9, 17)
    p = pl.compile_plan(tab, "dump")
    assert len(p.table_params) == 0
    kinds = [int(k) for k, _ in p.prefs]
    assert kinds.count(pl.PK_BETA_MUL) >= 10 and kinds.count(pl.PK_BETA) == 3
    W = pi.dump_plan(p, f)
    Y = orc.assemble(m.records(), m.n_nodes + len(m.sources), f)
    assert np.array_equal(W != 0, Y != 0)
    assert np.max(np.abs(W - Y)) <= 1e-13 * np.max(np.abs(Y))
    S = pi.run_plan(pl.compile_plan(tab), f)[0]
    So = orc.run_sp(m.records(), m.n_nodes, len(m.sources), f, method="lu_batched")
    assert np.max(np.abs(S - So)) < 1e-12


def test_shared_callable_is_tabulated_once():
    """One callable handed to several components is evaluated -- and uploaded -- once per run."""
    calls = []

    def z(f):
        calls.append(np.size(f))
        return 30.0 + 1e-9 * f

    m = hv.Model(suppress_loadbar=True)
    a, b = m.quick_port(50), m.quick_port(50)
    mid = m.node()
    m.impedance(a, mid, z)
    m.impedance(mid, b, z)
    m.impedance(mid, m.gnd, lambda f: 200.0 - 1j * 1e-8 * f)
    p = pl.compile_plan(m.stamp_table())
    assert len(p.table_params) == 3
    f = np.linspace(1e9, 2e9, 11)
    calls.clear()
    consts, prefs, tables, samples = pl.resolve_run(p, f)
    assert len(calls) == 1 and tables.shape == (2, 11)
    rows = [int(prefs[pi][1]) for (pi, _, _) in p.table_params]
    assert rows[0] == rows[1] != rows[2]


def test_band_ordering_with_two_port_blocks():
    """Two-port S blocks keep a cascade narrow-banded: the ordered plan is allowed (larger N-port blocks are
    not), keeps its cooperative ops, and solves to the same S as the natural order."""
    import circuits
    m, f = circuits.build_twoport_cascade(hv, sections=5, nF=7)
    tab = m.stamp_table()
    nat = pl.compile_plan(tab, order="natural")
    rcm = pl.compile_plan(tab, order="rcm")
    assert nat.band == -1 and 1 <= rcm.band <= 2 and len(rcm.coops) == len(nat.coops) == 5
    assert all(int(c[2]) == 2 for c in rcm.coops)
    assert np.max(np.abs(pi.run_plan(rcm, f) - pi.run_plan(nat, f))) < 1e-12
    # the 8-port Touchstone block of config 3 is never reordered
    assert pl.compile_plan(REG["c3_touchstone"]().stamp_table(), order="rcm").band == -1
