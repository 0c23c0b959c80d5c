"""GPU (-m gpu): edge cases of the CUDA path against the CPU oracle (LU variant, same float64
matrices): degenerate sweep lengths, every right-hand-side padding class, the kernel-class
boundaries, every parameter kind, and the state handling of repeated `run_sp` calls."""
import os

import numpy as np
import pytest

import heavi_b200 as hv
from heavi_b200 import plan as pl
from heavi_b200.engine import CompiledNetwork
from oracle import mna_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 2e-12


def oracle_S(m, f):
    return orc.run_sp(m.records(), m.n_nodes, len(m.sources), np.asarray(f, dtype=np.float64), method="lu_batched")


def ladder(n_sections, n_ports, z0=50.0, seed=0):
    """RLC ladder with `n_sections` internal nodes and `n_ports` ports spread along it."""
    rng = np.random.default_rng(seed)
    m = hv.Model(suppress_loadbar=True)
    nodes = [m.node() for _ in range(n_sections)]
    taps = sorted(set(int(round(i * (n_sections - 1) / max(n_ports - 1, 1))) for i in range(n_ports)))
    for t in taps:
        m.new_port(z0, nodes[t])
    for k in range(n_sections - 1):
        if k % 2:
            m.inductor(nodes[k], nodes[k + 1], float(10 ** rng.uniform(-9.3, -8.3)))
        else:
            m.resistor(nodes[k], nodes[k + 1], float(rng.uniform(5, 80)))
    for k in range(n_sections):
        m.capacitor(nodes[k], m.gnd, float(10 ** rng.uniform(-12.5, -11.5)))
    return m, len(taps)


@pytest.mark.parametrize("nF", [0, 1, 2, 3, 5, 31, 33, 127, 129])
def test_sweep_lengths(nF):
    m, _ = ladder(6, 2)
    f = np.linspace(0.5e9, 4e9, nF)
    S = m.run_sp(f)
    assert S._S.shape == (2, 2, nF) and S.f.shape == (nF,)
    if nF:
        assert np.max(np.abs(S._S - oracle_S(m, f))) < TOL


@pytest.mark.parametrize("n_sections,n_ports", [(1, 1), (3, 1), (4, 3), (6, 4), (9, 5), (12, 8), (14, 9), (20, 12),
                                               (30, 16), (24, 17), (31, 2), (32, 3), (33, 2), (40, 4), (70, 6)])
def test_shapes_and_kernel_classes(n_sections, n_ports, kernel_path):
    """P = 1..17 exercises every RHS padding class (2/4/8/16) and the P > 16 hand-over away from the
    register kernels; n = 31..33 crosses the register-kernel limit.  These ladders are tridiagonal, so
    what does not fit the register kernels goes to the banded kernel (class 2), not the dense block one."""
    m, P = ladder(n_sections, n_ports, seed=n_sections)
    f = np.array([0.3e9, 1.1e9, 2.7e9, 5.2e9, 7.9e9])
    S = m.run_sp(f)._S
    ref = oracle_S(m, f)
    assert S.shape == (P, P, 5)
    assert np.max(np.abs(S - ref)) < TOL
    ki = m._compiled().handle.kernel_info()
    plan = m._compiled().plan
    beyond_registers = plan.n_real > 32 or P > 16
    if kernel_path == "generated" and P <= 16 and plan.n_real >= 3:
        assert ki["kernel_class"] == 3            # a chain: the kernel generated for this netlist
    else:
        assert ki["kernel_class"] == (2 if beyond_registers else 0)


def test_full_mode_forced_and_block_forced_agree():
    m, P = ladder(10, 3, seed=5)
    f = np.linspace(0.2e9, 6e9, 77)
    ref = oracle_S(m, f)
    for mode in ("full", "norton"):
        for force in ("0", "1"):
            os.environ["HVB_FORCE_BLOCK"] = force
            try:
                S = CompiledNetwork(m, mode=mode).run(f)[0]
            finally:
                os.environ.pop("HVB_FORCE_BLOCK")
            assert np.max(np.abs(S - ref)) < TOL, (mode, force)


def test_every_parameter_kind_in_one_circuit():
    mc = hv.MonteCarlo()
    m = hv.Model(suppress_loadbar=True)
    a, b, c = m.quick_port(50), m.quick_port(75), m.quick_port(33.0)
    x, y = m.node(), m.node()
    Cr = mc.gaussian(1e-12, 5e-14)
    Lr = mc.uniform(1e-9, 2e-9)
    Lr._value = 1.5e-9            # a uniform Random is Uninitialized until drawn (numeric.py:233), and
    #                               Impedance.__init__ evaluates its value once (impedance.py:18)
    Gr = mc.gaussian(0.02, 0.001)                                          # a conductance, in siemens
    Rn = mc.gaussian(-40.0, 1.0)                                           # used negated: +40 ohm
    iL = mc.gaussian(5e8, 1e7)                                             # used inverted: 2 nH
    m.capacitor(a, m.gnd, Cr)                                              # SAMPLE
    m.inductor(a, x, Lr)                                                   # SAMPLE (uniform)
    m.admittance(x, m.gnd, Gr.inverse().inverse())                         # Inverse(Inverse(Random)) -> Random
    m.impedance(x, y, Rn.negative())                                       # SAMPLE_NEG
    m.capacitor(y, m.gnd, lambda f: 1e-12 * (1 + f / 1e10))                # TABLE (real)
    m.impedance(y, b, lambda f: 20 + 2j * np.pi * f * 1e-9)                # TABLE (complex)
    m.admittance(b, c, 0.01 + 0.002j)                                      # CONST complex
    m.resistor(c, m.gnd, 120.0)
    m.transmissionline(x, c, 60.0, complex(3.0, -0.05), 7e-3)              # closed-form beta, lossy
    m.TL(y, a, lambda f: 2 * np.pi * f / 299792458 * 1.5, 4e-3, 45.0)      # TABLE beta
    m.inductor(b, m.gnd, iL.inverse())                                     # SAMPLE_INV
    f = np.linspace(0.4e9, 5e9, 41)
    np.random.seed(5)
    vals = mc.draw(6)
    S4 = m.run_sp_batch(f, mc._random_numbers, vals)
    assert S4.shape == (6, 3, 3, 41)
    for s in range(6):
        for r, v in zip(mc._random_numbers, vals[s]):
            r._value = v
        assert np.max(np.abs(S4[s] - oracle_S(m, f))) < 5e-12, s
    kinds = {k for k, _ in m._compiled().plan.prefs.tolist()}
    assert {pl.PK_CONST, pl.PK_SAMPLE, pl.PK_SAMPLE_NEG, pl.PK_SAMPLE_INV, pl.PK_TABLE, pl.PK_BETA} <= kinds


def test_nport_y_block_equals_discrete_admittances():
    """The reference's NPortY cannot be constructed (App. B-4); ours must stamp the indefinite block."""
    y12 = lambda f: 1 / (30 + 2j * np.pi * f * 2e-9)
    y1g = lambda f: 2j * np.pi * f * 1e-12
    y2g = lambda f: 1 / 150.0 + 0 * f
    m1 = hv.Model(suppress_loadbar=True)
    a, b = m1.quick_port(50), m1.quick_port(50)
    Y = [[lambda f: y12(f) + y1g(f), lambda f: -y12(f)], [lambda f: -y12(f), lambda f: y12(f) + y2g(f)]]
    m1.n_port_Y(m1.gnd, [a, b], Y, 50.0)
    m1.resistor(a, m1.gnd, 1e6)
    m1.resistor(b, m1.gnd, 1e6)
    m2 = hv.Model(suppress_loadbar=True)
    a2, b2 = m2.quick_port(50), m2.quick_port(50)
    m2.admittance(a2, b2, y12)
    m2.admittance(a2, m2.gnd, y1g)
    m2.admittance(b2, m2.gnd, y2g)
    m2.resistor(a2, m2.gnd, 1e6)
    m2.resistor(b2, m2.gnd, 1e6)
    f = np.linspace(0.5e9, 6e9, 33)
    assert np.max(np.abs(m1.run_sp(f)._S - m2.run_sp(f)._S)) < 1e-12
    assert np.max(np.abs(m2.run_sp(f)._S - oracle_S(m2, f))) < TOL


def test_touchstone_two_port_ma_format_device_splines(tmp_path):
    fs = np.linspace(1.0, 4.0, 31)                       # GHz
    s11 = 0.3 * np.exp(-1j * fs)
    s21 = 0.8 * np.exp(-2j * fs)
    s12 = 0.1 * np.exp(-2j * fs)                          # non-reciprocal on purpose
    s22 = 0.2 * np.exp(0.5j * fs)
    path = tmp_path / "amp.s2p"
    with open(path, "w") as fh:
        fh.write("# GHz S MA R 50\n")
        for k in range(len(fs)):
            row = [fs[k]]
            for v in (s11[k], s21[k], s12[k], s22[k]):   # touchstone 2-port order
                row += [abs(v), np.degrees(np.angle(v))]
            fh.write(" ".join("%.12g" % x for x in row) + "\n")
    m = hv.Model(suppress_loadbar=True)
    a, b = m.quick_port(50), m.quick_port(50)
    n1, n2 = m.node(), m.node()
    m.resistor(a, n1, 5.0)
    m.capacitor(n2, m.gnd, 0.3e-12)
    m.inductor(n2, b, 0.7e-9)
    hv.FileBasedNPort(str(path)).connect(n1, n2)
    # inside, at, and outside the data range (constant end fill)
    f = np.concatenate([np.linspace(0.2e9, 5e9, 97), fs[:3] * 1e9, [1.0e9, 4.0e9]])
    plan = m._compiled().plan
    assert len(plan.spl_trace) == 4 and plan.n_tables == 0          # evaluated on the device, not tabulated
    assert np.max(np.abs(m.run_sp(f)._S - oracle_S(m, f))) < 5e-12


def test_repeated_runs_track_parameter_and_netlist_changes():
    sweep_param = hv.Param(np.array([1e-12, 2e-12, 4e-12]))
    m = hv.Model(suppress_loadbar=True)
    a, b = m.quick_port(50), m.quick_port(50)
    m.capacitor(a, m.gnd, sweep_param)
    m.inductor(a, b, 2e-9)
    f = np.linspace(1e9, 3e9, 64)
    outs = []
    for i in range(3):
        sweep_param.set_index(i)
        sweep_param.initialize()
        S = m.run_sp(f)._S.copy()
        assert np.max(np.abs(S - oracle_S(m, f))) < TOL
        outs.append(S)
    assert not np.array_equal(outs[0], outs[1])
    first = m._compiled()
    m.run_sp(f)
    assert m._compiled() is first                                    # plan reused while the netlist is unchanged
    m.resistor(b, m.gnd, 200.0)                                      # netlist edit after a run
    S = m.run_sp(f)._S
    assert m._compiled() is not first
    assert np.max(np.abs(S - oracle_S(m, f))) < TOL


def test_unsorted_duplicate_and_log_spaced_frequencies():
    m, _ = ladder(7, 2, seed=3)
    f = np.concatenate([np.logspace(6, 10.3, 50)[::-1], [2e9, 2e9, 1e6]])
    S = m.run_sp(f)._S
    assert np.max(np.abs(S - oracle_S(m, f))) < TOL
    assert np.array_equal(S[:, :, 50], S[:, :, 51])


def test_badly_scaled_admittances():
    m = hv.Model(suppress_loadbar=True)
    a, b = m.quick_port(50), m.quick_port(50)
    x = m.node()
    m.resistor(a, x, 1e-3)
    m.resistor(x, b, 1e5)
    m.capacitor(x, m.gnd, 1e-15)
    m.inductor(x, m.gnd, 1e-3)
    f = np.array([1e3, 1e6, 1e9, 1e11])
    S = m.run_sp(f)._S
    ref = oracle_S(m, f)
    assert np.max(np.abs(S - ref)) < 1e-11


def test_pageable_output_buffer_and_status_word():
    m, _ = ladder(5, 2, seed=9)
    cn = m._compiled()
    f = np.linspace(1e9, 2e9, 1000)
    consts, prefs, tables, samples = cn.resolve(f)
    out = np.empty((1, 2, 2, 1000), dtype=np.complex128)             # ordinary (pageable) numpy memory
    S, status = cn.handle.run_host(f, consts, prefs, tables, samples, out)
    assert S is out and status == 0
    assert np.max(np.abs(out[0] - oracle_S(m, f))) < TOL


def test_two_models_interleaved():
    m1, _ = ladder(5, 2, seed=1)
    m2, _ = ladder(11, 3, seed=2)
    f = np.linspace(1e9, 5e9, 200)
    for _ in range(3):
        assert np.max(np.abs(m1.run_sp(f)._S - oracle_S(m1, f))) < TOL
        assert np.max(np.abs(m2.run_sp(f)._S - oracle_S(m2, f))) < TOL


def test_frequency_axis_float32_filled_by_the_library():
    """`Sparameters.f` is float32 (network.py:407,421); the library converts it while the GPU works."""
    m, _ = ladder(6, 2)
    f = np.linspace(0.5e9, 4e9, 1001) + 0.123
    S = m.run_sp(f)
    assert S.f.dtype == np.float32 and np.array_equal(S.f, f.astype(np.float32))
    S2 = m.run_sp(list(f[:7]))
    assert np.array_equal(S2.f, f[:7].astype(np.float32))


def test_monte_carlo_with_sample_dependent_callables_runs_per_trial():
    """A closure over Random objects (lib/mc.py RandomTwoPort) changes its table from trial to trial: the
    batched drivers must notice and reproduce the reference's serial loop, seed for seed."""
    import circuits
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "mc_twoport.npz"))
    m, f, mc = circuits.build_random_twoport(hv)
    cn = m._compiled()
    np.random.seed(11)
    vals = mc.draw(6)
    assert cn._tables_follow_drivers(f, mc._random_numbers, vals)
    np.random.seed(11)
    stat = m.run_mc(f, mc, 6)
    S = np.stack(stat._Sdata) if isinstance(stat._Sdata, list) else np.asarray(stat._Sdata)
    assert S.shape == g["mc_S"].shape and np.max(np.abs(S - g["mc_S"])) < 1e-11
    np.random.seed(11)
    mom = m.run_mc_stats(f, mc, 6)
    for (j, i) in ((1, 1), (2, 1), (1, 2)):
        ref = np.abs(g["mc_S"][:, j - 1, i - 1, :])
        assert np.max(np.abs(mom.S(j, i).amplitude_mean - ref.mean(axis=0))) < 1e-11
        assert np.max(np.abs(mom.S(j, i).amplitude_std - ref.std(axis=0))) < 1e-11
    # plain Random-valued components (config 4) stay on the batched path
    m4, f4, mc4 = circuits.build_c4(hv)
    np.random.seed(7)
    assert not m4._compiled()._tables_follow_drivers(f4, mc4._random_numbers, mc4.draw(3))


def test_smd_parts_are_evaluated_on_the_device():
    """Surface-mount models (resistor, inductor, capacitor with package parasitics) are closed forms on the
    device: assembled stamps within a few ulp of the host's numpy evaluation, S equal to the oracle."""
    from test_plan import _smd_model
    from helpers import max_ulps
    m = _smd_model()
    cn = m._compiled()
    assert len(cn.plan.table_params) == 0
    f = np.linspace(0.3e9, 7e9, 57)
    Y = cn.assemble(f)
    Yo = orc.assemble(m.records(), m.n_nodes + len(m.sources), f)
    assert np.array_equal(Y != 0, Yo != 0) and max_ulps(Y, Yo) <= 4.0
    assert np.max(np.abs(m.run_sp(f)._S - oracle_S(m, f))) < TOL


def test_coupler_library_on_the_device():
    """Wilkinson / branch-line / rat-race / Chebyshev transformer (lib/tl.py): phase constants evaluated on
    the device in the library's own operation order."""
    from test_plan import _coupler_model
    m = _coupler_model()
    cn = m._compiled()
    assert len(cn.plan.table_params) == 0
    f = np.linspace(1.2e9, 3.6e9, 41)
    Y = cn.assemble(f)
    Yo = orc.assemble(m.records(), m.n_nodes + len(m.sources), f)
    assert np.array_equal(Y != 0, Yo != 0) and np.max(np.abs(Y - Yo)) <= 1e-13 * np.max(np.abs(Yo))
    assert np.max(np.abs(m.run_sp(f)._S - oracle_S(m, f))) < TOL


def test_resident_results_match_the_host_path():
    """run_sp_resident leaves S in HBM (torch CUDA tensor); same values as run_sp / run_sp_batch."""
    import torch
    import circuits
    m, _ = ladder(9, 3, seed=2)
    f = np.linspace(0.4e9, 5e9, 777)
    S_dev = m.run_sp_resident(f)
    assert S_dev.is_cuda and S_dev.dtype == torch.complex128 and tuple(S_dev.shape) == (1, 3, 3, 777)
    assert np.array_equal(S_dev.cpu().numpy()[0], m.run_sp(f)._S)
    S_dev2 = m.run_sp_resident(torch.from_numpy(f).cuda())               # frequencies already on the device
    assert torch.equal(S_dev, S_dev2)
    m4, f4, mc = circuits.build_c4(hv, nF=201)
    np.random.seed(3)
    vals = mc.draw(16)
    Sb = m4.run_sp_resident(f4, mc._random_numbers, vals)
    assert np.array_equal(Sb.cpu().numpy(), m4.run_sp_batch(f4, mc._random_numbers, vals))
    # a derived quantity computed where the data lives: worst-case |S21| in dB over the samples
    worst = (20 * torch.log10(Sb[:, 1, 0, :].abs())).amin(dim=0)
    assert worst.shape == (201,) and bool(torch.isfinite(worst).all())
    m0 = hv.Model(suppress_loadbar=True)
    a, b = m0.quick_port(50), m0.quick_port(50)
    m0.inductor(a, b, 1e-9)
    with pytest.raises(np.linalg.LinAlgError):
        m0.run_sp_resident(np.array([0.0, 1e9]))


def test_line_length_follows_its_parameter():
    """The reference reads `self.length.value` inside every run_sp (base/tl.py:52): a swept / drawn /
    optimised line length must change S from run to run and from sample to sample."""
    m = hv.Model(suppress_loadbar=True)
    sweep = hv.ParameterSweep()
    Ln = sweep.lin(0.004, 0.031).add(5)
    Ln.initialize()
    a, b = m.quick_port(50), m.quick_port(50)
    mid = m.node()
    m.transmissionline(a, mid, 63.0, 2.9, Ln)
    m.capacitor(mid, m.gnd, 0.7e-12)
    m.inductor(mid, b, 1.5e-9)
    f = np.linspace(0.8e9, 5e9, 64)
    nd = m.run_sweep(f, sweep)
    assert nd._sdata.shape == (5, 2, 2, 64)
    for k, v in enumerate(Ln._values):
        Ln._value = v
        ref = oracle_S(m, f)
        assert np.max(np.abs(nd._sdata[k] - ref)) < TOL
        assert np.max(np.abs(m.run_sp(f)._S - ref)) < TOL          # the serial loop sees the new length too
    assert np.max(np.abs(nd._sdata[0] - nd._sdata[4])) > 1e-3


def test_run_sp_many_overlaps_independent_models_and_equals_serial_calls():
    import circuits
    jobs = [(circuits.build_c2(hv, band, nF=8)[0], np.linspace(2e9, 3e9, 30_001)) for band in ("bandpass", "lowpass", "highpass")]
    jobs.append((circuits.build_c1(hv, nF=8)[0], np.linspace(1.8e9, 2.2e9, 2001)))
    many = hv.run_sp_many(jobs)
    for (m, f), S in zip(jobs, many):
        serial = m.run_sp(f)
        assert np.array_equal(S._S, serial._S) and np.array_equal(S.f, serial.f) and S.f.dtype == np.float32
    # one run in flight per plan
    m, f = jobs[0]
    p = m.run_sp_async(f)
    with pytest.raises(ValueError):
        m.run_sp(f)
    assert np.array_equal(p.result()._S, many[0]._S)
    assert np.array_equal(m.run_sp(f)._S, many[0]._S)
