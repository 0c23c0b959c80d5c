"""Kernel C (banded LU, one thread per point) against the dense kernels, the oracle and the golden
ladder.  The banded path is chosen by the library when the RCM-ordered system is narrow-banded
(heavi_b200/csrc/hvb_band.cu); HVB_NO_BAND=1 forces the dense block kernel for the same plan."""
import os

import numpy as np
import pytest

import circuits
import helpers

pytestmark = [pytest.mark.gpu, pytest.mark.library_kernels]

TOL_REL, TOL_ABS = 1e-10, 1e-12


def _compile(hv, m, band: bool, mode=None):
    from heavi_b200.engine import CompiledNetwork
    old = os.environ.get("HVB_NO_BAND")
    os.environ["HVB_NO_BAND"] = "0" if band else "1"
    try:
        return CompiledNetwork(m, mode=mode)
    finally:
        if old is None:
            os.environ.pop("HVB_NO_BAND", None)
        else:
            os.environ["HVB_NO_BAND"] = old


def _close(a, b):
    return np.all(np.abs(a - b) <= TOL_ABS + TOL_REL * np.maximum(np.abs(a), np.abs(b)))


def test_c5_uses_banded_kernel_and_matches_golden():
    import heavi_b200 as hv
    m, _ = circuits.build_c5(hv, nF=10)
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "c5_ladder.npz"))
    cn = _compile(hv, m, band=True)
    ki = cn.handle.kernel_info()
    assert ki["kernel_class"] == 2 and ki["group"] == 1 and cn.plan.band == 1
    S = cn.run(g["f"])[0]
    assert _close(S, g["S_ext"])                      # 80-bit adjudicator
    dense = _compile(hv, m, band=False)
    assert dense.handle.kernel_info()["kernel_class"] == 1
    assert _close(S, dense.run(g["f"])[0])


@pytest.mark.parametrize("reach,K,nports,mode", [(1, 40, 2, None), (2, 60, 3, None), (3, 64, 4, None), (4, 80, 3, None),
                                                  (5, 90, 2, None), (7, 120, 5, None), (8, 150, 8, None),
                                                  (2, 50, 3, "full"), (1, 33, 2, "full"), (4, 70, 16, None)])
def test_banded_matches_dense_and_oracle(reach, K, nports, mode):
    import heavi_b200 as hv
    from oracle import mna_oracle as orc
    m, f = circuits.build_band_ladder(hv, K=K, reach=reach, nports=nports, seed=reach * 100 + K)
    cn = _compile(hv, m, band=True, mode=mode)
    ki = cn.handle.kernel_info()
    assert ki["kernel_class"] == 2, ki
    assert cn.plan.band <= ki["group"]
    S = cn.run(f)[0]
    dense = _compile(hv, m, band=False, mode=mode)
    assert dense.handle.kernel_info()["kernel_class"] == 1
    Sd = dense.run(f)[0]
    assert _close(S, Sd)
    So = orc.run_sp(m.records(), m.n_nodes, len(m.sources), np.asarray(f, dtype=np.float64), method="lu_batched")
    assert _close(S, So)


def test_banded_many_points_are_consistent():
    """More points than threads in flight for a small grid: every thread reuses its workspace."""
    import heavi_b200 as hv
    m, _ = circuits.build_band_ladder(hv, K=48, reach=2, nports=2, seed=7)
    f = np.linspace(0.5e9, 6e9, 70_001)
    cn = _compile(hv, m, band=True)
    S = cn.run(f)[0]
    idx = np.array([0, 1, 127, 128, 4095, 35_000, 69_999, 70_000])
    S_few = cn.run(f[idx])[0]
    assert np.array_equal(S[..., idx], S_few)
    dense = _compile(hv, m, band=False)
    assert _close(S[..., ::997], dense.run(f[::997])[0])


def test_banded_monte_carlo_samples():
    import heavi_b200 as hv
    m = hv.Model()
    mc = hv.MonteCarlo()
    chain = [m.node() for _ in range(41)]
    m.new_port(50, chain[0]); m.new_port(50, chain[40])
    for k in range(40):
        if k % 2:
            m.inductor(chain[k], chain[k + 1], mc.gaussian(2e-9, 1e-10))
        else:
            m.transmissionline(chain[k], chain[k + 1], 50.0, 2.9, 4e-3)
        m.capacitor(chain[k + 1], m.gnd, mc.gaussian(1e-12, 5e-14))
    f = np.linspace(1e9, 3e9, 33)
    np.random.seed(5)
    values = mc.draw(24)
    Sb = _compile(hv, m, band=True).run(f, mc._random_numbers, values)
    Sd = _compile(hv, m, band=False).run(f, mc._random_numbers, values)
    assert Sb.shape == (24, 2, 2, 33)
    assert _close(Sb, Sd)
    assert not np.array_equal(Sb[0], Sb[1])


def test_banded_singular_point_is_flagged():
    import heavi_b200 as hv
    m = hv.Model()
    chain = [m.node() for _ in range(40)]
    m.new_port(50, chain[0]); m.new_port(50, chain[39])
    for k in range(39):
        if k != 20:                                   # node 20..39 side floats from the rest: still fine
            m.resistor(chain[k], chain[k + 1], 10.0)
    floating = m.node()                               # a node nothing else touches but a zero admittance pair
    m.admittance(floating, chain[5], 0.0)
    import warnings
    cn = _compile(hv, m, band=True)
    assert cn.handle.kernel_info()["kernel_class"] == 2
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        try:
            cn.run(np.array([1e9, 2e9]))
            flagged = any("singular" in str(x.message) for x in w)
        except np.linalg.LinAlgError:
            flagged = True
    assert flagged


@pytest.mark.parametrize("band", [True, False])
def test_two_pipeline_slots_do_not_share_a_workspace(band):
    """hvb_run_host alternates two streams; kernels B and C keep their matrices in a global-memory
    workspace, so each stream needs its own or the tail of one tile corrupts the head of the next."""
    import heavi_b200 as hv
    m, _ = circuits.build_c5(hv, nF=10)
    f = np.linspace(0.5e9, 6e9, 140_000)
    cn = _compile(hv, m, band=band)
    os.environ["HVB_CHUNK_MB"] = "1"
    try:
        S = cn.run(f)[0].copy()
    finally:
        os.environ.pop("HVB_CHUNK_MB")
    idx = np.unique(np.concatenate([np.arange(0, 140_000, 1511), np.arange(65_400, 65_700), np.arange(69_900, 70_100)]))
    assert np.array_equal(S[..., idx], cn.run(f[idx])[0])


def test_mid_size_banded_system_switches_kernel_with_the_batch_size():
    """9..32 unknowns, narrow band: short sweeps stay on the register kernel (lowest latency), large
    batches go to the banded kernel (one thread per point: 2-9x the throughput); same S either way."""
    import heavi_b200 as hv
    from heavi_b200.engine import CompiledNetwork
    m, _ = circuits.build_band_ladder(hv, K=19, reach=1, nports=2, seed=21)
    cn = CompiledNetwork(m)
    assert cn.plan.n_real == 20 and cn.handle.kernel_info()["kernel_class"] == 0
    f_small = np.linspace(0.5e9, 6e9, 501)
    S_small = cn.run(f_small)[0]
    assert cn._band is None                                   # banded variant not even built yet
    f_big = np.linspace(0.5e9, 6e9, CompiledNetwork.band_min_points(20) + 17)
    S_big = cn.run(f_big)[0]
    assert cn._band and cn._band[1].kernel_info()["kernel_class"] == 2 and cn._band[0].band == 1
    # the same frequencies through both kernels
    idx = np.linspace(0, len(f_big) - 1, 301).astype(int)
    S_dense = cn.handle.run_host(f_big[idx], *cn.resolve(f_big[idx]))[0][0]
    assert _close(S_big[..., idx], S_dense)
    from oracle import mna_oracle as orc
    So = orc.run_sp(m.records(), m.n_nodes, len(m.sources), f_small, method="lu_batched")
    assert _close(S_small, So)
    # Monte-Carlo batches count samples x frequencies
    mc = hv.MonteCarlo()
    m2 = hv.Model(suppress_loadbar=True)
    chain = [m2.node() for _ in range(12)]
    m2.new_port(50, chain[0]); m2.new_port(50, chain[11])
    for k in range(11):
        m2.inductor(chain[k], chain[k + 1], mc.gaussian(2e-9, 1e-10))
        m2.capacitor(chain[k + 1], m2.gnd, mc.gaussian(1e-12, 5e-14))
    f = np.linspace(1e9, 3e9, 257)
    np.random.seed(3)
    vals = mc.draw(200)                                         # 51 400 points >= band_min_points(11)
    cn2 = m2._compiled()
    Sb = cn2.run(f, mc._random_numbers, vals)
    assert cn2._band and Sb.shape == (200, 2, 2, 257)
    Sd = cn2.handle.run_host(f, *cn2.resolve(f, mc._random_numbers, vals[:5]))[0]
    assert _close(Sb[:5], Sd)


def test_sharded_run_of_a_mid_size_banded_system_is_bit_identical():
    """Each shard of a distributed run is told the size of the whole job, so all of them use the kernel
    variant the single-GPU run uses and the pieces concatenate to exactly the same S."""
    import heavi_b200 as hv
    from heavi_b200 import sharding as sh
    m, _ = circuits.build_band_ladder(hv, K=15, reach=1, nports=2, seed=4)
    cn = m._compiled()
    f = np.linspace(0.5e9, 6e9, cn.band_min_points(cn.plan.n_real) + 1000)
    whole = cn.run(f)[0].copy()
    assert cn._band                                            # the whole job is large enough for the banded kernel
    for world in (2, 8):                                       # each shard alone would be below the threshold
        parts = [sh.run_sharded(lambda fb, vb: cn.run(fb, batch_points=len(f)).copy(), f, None, world, r) for r in range(world)]
        assert np.array_equal(np.concatenate([p[1] for p in parts], axis=3)[0], whole)


def test_cascade_of_two_port_blocks_runs_on_the_banded_kernel():
    """Two-port S blocks are evaluated per thread inside the banded kernel (larger N-port blocks keep the
    lane-cooperative register kernels): same S as the dense path and as the oracle."""
    import heavi_b200 as hv
    from heavi_b200.engine import CompiledNetwork
    from oracle import mna_oracle as orc
    m, f = circuits.build_twoport_cascade(hv, sections=6)
    cn = CompiledNetwork(m)
    assert cn.plan.n_real >= 12 and cn.handle.kernel_info()["kernel_class"] == 0 and len(cn.plan.coops) == 6
    band = cn._band_variant()
    assert band and band[1].kernel_info()["kernel_class"] == 2 and band[0].band <= 2
    Sb = band[1].run_host(f, *cn.resolve(f, plan=band[0]))[0][0]
    Sd = cn.run(f)[0]
    assert _close(Sb, Sd)
    So = orc.run_sp(m.records(), m.n_nodes, len(m.sources), f, method="lu_batched")
    assert _close(Sb, So)
    # a long cascade (beyond the register kernels) takes the banded kernel by default
    m2, f2 = circuits.build_twoport_cascade(hv, sections=20, seed=3)
    cn2 = CompiledNetwork(m2)
    assert cn2.plan.n_real > 32 and cn2.handle.kernel_info()["kernel_class"] == 2
    S2 = cn2.run(f2)[0]
    So2 = orc.run_sp(m2.records(), m2.n_nodes, len(m2.sources), f2, method="lu_batched")
    assert _close(S2, So2)
