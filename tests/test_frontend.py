"""CPU: behaviour of the drop-in building API (reference: heavi/network.py, node.py, sparam.py)."""
import numpy as np
import pytest

import heavi_b200 as hv
from heavi_b200.results import Sparameters


def test_port_creates_internal_node_before_signal_node():
    m = hv.Model()
    p = m.new_port(50)
    m.resistor(p.sig_node, m.gnd, 50.0)
    m._prepare()
    assert m.gnd.index == 0 and p.int_node.index == 1 and p.sig_node.index == 2
    assert p.sources[0].index == 3 and p.index == 1


def test_unconnected_node_raises_value_error():
    m = hv.Model()
    a = m.quick_port(50)
    m.node()                                       # dangling
    m.resistor(a, m.gnd, 10.0)
    with pytest.raises(ValueError, match="is not connected to any components"):
        m.stamp_table()


def test_merged_nodes_share_an_index():
    m = hv.Model()
    a = m.quick_port(50)
    b = m.node()
    c = m.node()
    assert (c > b) is b
    m.resistor(a, b, 10.0)
    m.capacitor(c, m.gnd, 1e-12)
    m._prepare()
    assert c.index == b.index
    assert max(n.index for n in m._nodes) + 1 == 4


def test_quick_port_forwards_second_argument_as_signal_node():
    m = hv.Model()
    n = m.node()
    assert m.quick_port(50, n) is n                # reference quirk, network.py:500


def test_reinitialize_does_not_grow_sources():
    m = hv.Model()
    a = m.quick_port(50)
    m.resistor(a, m.gnd, 75.0)
    m._prepare()
    m._initialized = False
    m._prepare()
    assert len(m.sources) == 1


def test_sparameters_container():
    S = np.arange(2 * 2 * 3, dtype=np.complex128).reshape(2, 2, 3)
    sp = Sparameters(S, np.array([1, 2, 3], dtype=np.float32))
    assert sp.nports == 2 and sp.nfreqs == 3 and sp.shape == (2, 2, 3)
    assert np.array_equal(sp.S21, S[1, 0]) and np.array_equal(sp(1, 2), S[0, 1]) and np.array_equal(sp.S(2, 2), S[1, 1])
    with pytest.raises(ValueError):
        sp.S(3, 1)
    with pytest.raises(ValueError):
        sp.S110                                    # greedy regex -> (11, 0), like the reference
    with pytest.raises(AttributeError):
        sp.nothing
    assert "S54" in dir(sp) and "S55" not in [x for x in dir(sp)[:20]]


def test_scalar_param_semantics():
    from heavi_b200.params import Scalar, enforce_simparam, Function, Random
    s = Scalar(3.0)
    assert np.array_equal(s(np.array([1.0, 2.0])), np.array([3.0, 3.0])) and s.value == 3.0
    assert isinstance(enforce_simparam(lambda f: f), Function)
    with pytest.raises(ValueError):
        enforce_simparam("x")
    mc = hv.MonteCarlo()
    u = mc.uniform(0, 1)
    with pytest.raises(ValueError, match="Uninitialized"):
        u(np.array([1.0]))
    assert isinstance(u, Random)


def test_touchstone_roundtrip(tmp_path):
    f = np.linspace(1e9, 2e9, 7)
    S = np.zeros((7, 2, 2), dtype=complex)
    S[:, 0, 0], S[:, 1, 0], S[:, 0, 1], S[:, 1, 1] = 0.1 + 0.2j, 0.8 - 0.1j, 0.05j, 0.3
    path = tmp_path / "x.s2p"
    with open(path, "w") as fh:
        fh.write("! test\n# GHz S RI R 75\n")
        for k in range(7):
            # 2-port order: S11 S21 S12 S22
            vals = [S[k, 0, 0], S[k, 1, 0], S[k, 0, 1], S[k, 1, 1]]
            fh.write("%.9f " % (f[k] / 1e9) + " ".join("%.6f %.6f" % (v.real, v.imag) for v in vals) + "\n")
    blk = hv.FileBasedNPort(str(path))
    assert blk.n_ports == 2 and blk.refimp == 75.0 and blk.freq_unit == 1e9
    assert np.allclose(blk.Sdata, S) and np.allclose(blk.Fdata, f)
    with pytest.raises(ValueError):
        hv.FileBasedNPort(str(tmp_path / "missing.s2p"))
    bad = tmp_path / "x.s8p"
    bad.write_text(path.read_text())
    with pytest.raises(ValueError):
        hv.FileBasedNPort(str(bad))                # suffix rejected unless ignore_extension
    assert hv.FileBasedNPort(str(bad), ignore_extension=True).n_ports == 2


def test_cauer_filter_values_and_counts():
    m = hv.Model()
    a, b = m.quick_port(50), m.quick_port(50)
    caps, inds, load = m.filters.cauer_filter(m.gnd, a, b, 2.5e9, 0.2e9, 5, 0.05, hv.FilterType.CHEBYCHEV,
                                              hv.BandType.BANDPASS, 50)
    assert len(caps) == 5 and len(inds) == 5 and load == pytest.approx(1.0)
    # series branch resonates at the centre frequency
    assert 1 / (2 * np.pi * np.sqrt(caps[0].value * inds[0].value)) == pytest.approx(2.5e9, rel=1e-12)
    m._prepare()
    assert m.n_nodes == 10 and len(m.sources) == 2


def test_run_sp_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = hv.Model()
    a = m.quick_port(50)
    m.resistor(a, m.gnd, 10.0)
    with pytest.raises(RuntimeError):
        m.run_sp(np.array([1e9]))


def test_parameter_sweep_combinations_and_cap():
    sweep = hv.ParameterSweep()
    a = sweep.lin(1.0, 2.0).add(3)
    b, c = sweep.lin(0.0, 1.0).lin(5.0, 6.0).add(2)
    assert [p._values.tolist() for p in sweep.params] == [[1.0, 1.5, 2.0], [0.0, 1.0], [5.0, 6.0]]
    assert sweep.combinations() == [(0, 0), (0, 1), (1, 0), (1, 1), (2, 0), (2, 1)]
    seen = [vals for _, vals in sweep.iterate()]
    assert seen[1] == (1.0, 1.0, 6.0) and len(seen) == 6
    # quirk kept from the reference (numeric.py:423-429): the 10 000 cap multiplies the number of
    # *parameters* per dimension, not the number of values, so a 101 x 100 grid does not trip it
    big = hv.ParameterSweep()
    big.lin(0, 1).add(101)
    big.lin(0, 1).add(100)
    assert len(big.combinations()) == 10_100


def test_optimiser_metrics_and_variables():
    from heavi_b200.lib import optim
    v = optim.OptimVar(-12.0, bounds=(-13, -6), mapping=lambda x: 10 ** x)
    assert v(np.array([1.0, 2.0])).tolist() == [1e-12, 1e-12] and v.describe()[0] == "sample"
    v.set(-11.0)
    assert v._value == pytest.approx(1e-11)
    with pytest.raises(TypeError):
        optim.OptimVar(1.0, mapping=3)
    S = np.array([0.1, 0.5, 1.0])
    assert optim.dBbelow(-10)(S) == pytest.approx(np.sqrt(np.mean(np.clip(20 * np.log10(S) + 10, 0, None) ** 2)) / 10)
    assert optim.dBabove(-3, optim.PNorm.MAX)(S) == pytest.approx(np.max(np.clip(-3 - 20 * np.log10(S), 0, None)) / 10)
    with pytest.raises(ValueError):
        optim.FrequencyLevel(1e9, 2e9, 5, (1, 1))
    m = hv.Model()
    opt = optim.Optimiser(m)
    c, l = opt.cap(), opt.ind(logscale=False)
    assert opt.bounds == [(-13, -6), (1e-12, 1e-5)] and opt.x0.tolist() == [-12.0, 1e-9]


def test_subcircuit_framework():
    from heavi_b200.lib.libgen import SubCircuit, TwoNodeSubCircuit, ThreeNodeSubCircuit, FourNodeSubCircuit

    class Pad(TwoNodeSubCircuit):
        def __on_connect__(self):
            self.network.resistor(self.node(1), self.node(2), 10.0)

    m = hv.Model()
    a, b, c = m.quick_port(50), m.node(), m.quick_port(50)
    p1 = Pad()
    assert (p1 < a) is p1 and (p1 > b) is b              # attach 1, attach 2 -> expands, returns the node
    assert len([x for x in m.components if x.kind == "resistor"]) == 1
    Pad().connect(b, c)
    assert len([x for x in m.components if x.kind == "resistor"]) == 2
    with pytest.raises(NotImplementedError):
        SubCircuit().__on_connect__()
    with pytest.raises(ValueError, match="at least one node"):
        s = SubCircuit(); s.network = m; s.connect()
    other = hv.Model()
    with pytest.raises(ValueError, match="same Network"):
        Pad().connect(a, other.node())
    assert ThreeNodeSubCircuit().n_nodes == 3 and FourNodeSubCircuit().n_nodes == 4

    class Tee(ThreeNodeSubCircuit):
        def __on_connect__(self):
            for i in (1, 2, 3):
                self.network.resistor(self.node(i), self.network.gnd, 100.0)

    t = Tee()
    t.t(1) > a
    t.t(3) < c
    assert t.first_node == 2 and not t._complete()
    t.partial_connect(2, b)
    assert len([x for x in m.components if x.kind == "resistor"]) == 5


def test_emerge_model_equals_touchstone_block():
    """EMergeModel((F, S)) is the in-memory counterpart of FileBasedNPort (lib/emerge.py:14-60): fed the
    arrays a Touchstone file parses to, it must produce the same netlist, plan and S."""
    import circuits
    import plan_interp as pi
    from heavi_b200 import plan as pl
    from heavi_b200.lib.emerge import EMergeModel
    ref_blk = hv.FileBasedNPort(circuits.SYNTH8)

    def build(block):
        m = hv.Model(suppress_loadbar=True)
        inner = [m.node() for _ in range(8)]
        for k, n in enumerate(inner):
            sig = m.quick_port(50)
            mid = m.node()
            m.inductor(sig, mid, (1.0 + 0.25 * k) * 1e-9)
            m.capacitor(mid, m.gnd, (0.5 + 0.1 * k) * 1e-12)
            m.capacitor(mid, n, (2.0 + 0.3 * k) * 1e-12)
        block.connect(*inner)
        return m

    m_file = build(hv.FileBasedNPort(circuits.SYNTH8))
    m_mem = build(EMergeModel((ref_blk.Fdata, ref_blk.Sdata)))
    ta, tb = m_file.stamp_table(), m_mem.stamp_table()
    assert ta.kinds == tb.kinds and ta.ids == tb.ids
    f = np.linspace(1e9, 6e9, 9)
    Sa = pi.run_plan(pl.compile_plan(ta), f)
    Sb = pi.run_plan(pl.compile_plan(tb), f)
    assert np.array_equal(Sa, Sb)
    blk = EMergeModel((ref_blk.Fdata, ref_blk.Sdata.copy()))
    assert blk.n_ports == 8 and blk.refimp == 50
    S0 = blk.Sdata.copy()
    blk.renormalize(75.0).renormalize(50.0)
    assert blk.refimp == 50.0 and np.max(np.abs(blk.Sdata - S0)) < 1e-12
    with pytest.raises(ValueError):
        EMergeModel((ref_blk.Fdata[:-1], ref_blk.Sdata))


def test_random_twoport_draws_and_per_trial_plan_match_the_reference():
    """lib/mc.py RandomTwoPort: the S_ij are closures over Random objects.  The pre-drawn batch must walk
    np.random exactly like `mc.iterate`, and a per-trial evaluation must reproduce the reference's S."""
    import os
    import circuits
    import plan_interp as pi
    from heavi_b200 import plan as pl
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "mc_twoport.npz"))
    m, f, mc = circuits.build_random_twoport(hv)
    np.random.seed(11)
    draws = mc.draw(6)
    assert np.array_equal(draws, g["mc_draws"])
    plan = pl.compile_plan(m.stamp_table())
    assert len(plan.table_params) == 4                           # the four S_ij closures are host-tabulated
    for s in range(6):
        for r, v in zip(mc._random_numbers, draws[s]):
            r._value = float(v)
        S = pi.run_plan(plan, f, mc._random_numbers, draws[s:s + 1])[0]
        assert np.max(np.abs(S - g["mc_S"][s])) < 1e-11
