"""CPU: the numpy oracle against the golden vectors generated from the REAL reference
(oracle/make_golden.py), driven through the product front-end's records -- so this also pins the
front-end's node/source numbering, component order and element values."""
import os
import numpy as np
import pytest

import heavi_b200 as hv
from helpers import dense_y, golden, registry
from oracle import mna_oracle as orc

REG = registry(hv)


@pytest.mark.parametrize("name", sorted(REG))
def test_stamp_ids_exact(name):
    g = golden(name)
    m = REG[name]()
    kinds, off, ids = m.stamp_table().flat()
    assert list(kinds) == list(g["kinds"])
    assert np.array_equal(off, g["off"])
    assert np.array_equal(ids, g["ids"])                      # bit-exact node / source indexing
    tab = m.stamp_table()
    assert np.array_equal(tab.port_table, g["indices"]) and tab.port_table.dtype == np.int32
    assert (tab.N, tab.M) == (int(g["N"]), int(g["M"]))


@pytest.mark.parametrize("name", sorted(REG))
def test_oracle_assembly_bit_exact(name):
    g = golden(name)
    m = REG[name]()
    Y = orc.assemble(m.records(), int(g["N"]) + int(g["M"]), g["y_f"])
    assert np.array_equal(Y, dense_y(g))                      # same numpy arithmetic as the reference


@pytest.mark.parametrize("name", [n for n in sorted(REG) if n not in ("c5_ladder",)])
def test_oracle_lstsq_matches_reference(name):
    g = golden(name)
    m = REG[name]()
    f = g["f"][:256]
    S = orc.run_sp(m.records(), int(g["N"]), int(g["M"]), f, method="lstsq")
    # same LAPACK driver (zgelsd) as the reference's numba lstsq: agreement far below 1e-12
    assert np.max(np.abs(S - g["S"][:, :, :len(f)])) < 5e-14


def test_oracle_lstsq_c5_subset():
    g = golden("c5_ladder")
    m = REG["c5_ladder"]()
    S = orc.run_sp(m.records(), int(g["N"]), int(g["M"]), g["f"][:2], method="lstsq")
    assert np.max(np.abs(S - g["S"][:, :, :2])) < 5e-14


@pytest.mark.parametrize("name", ["c1_readme", "c2_bandpass", "c3_touchstone", "c5_ladder", "rnd_07", "rnd_30"])
def test_oracle_lu_close_to_extended_precision(name):
    g = golden(name)
    m = REG[name]()
    S = orc.run_sp(m.records(), int(g["N"]), int(g["M"]), g["f"], method="lu_batched")
    assert np.max(np.abs(S - g["S_ext"])) < 5e-14
    # and the reference's own distance from the exact answer is what bounds |S_lu - S_ref|
    assert np.max(np.abs(S - g["S"])) <= np.max(np.abs(g["S"] - g["S_ext"])) + 5e-14


def test_monte_carlo_draw_order_matches_reference_loop():
    """`MonteCarlo.draw(N)` must consume np.random exactly like `iterate(N)` (numeric.py:498-503)."""
    import circuits
    g = golden("c4_mc")
    m, f, mc = circuits.build_c4(hv)
    np.random.seed(7)
    vals = mc.draw(4)
    assert np.array_equal(vals, g["mc_draws"])
    np.random.seed(7)
    seq = []
    for _ in mc.iterate(4):
        seq.append([r._value for r in mc._random_numbers])
    assert np.array_equal(np.array(seq), g["mc_draws"])


def test_oracle_mc_samples():
    import circuits
    g = golden("c4_mc")
    m, f, mc = circuits.build_c4(hv)
    for s in range(2):
        for k, r in enumerate(mc._random_numbers):
            r._value = g["mc_draws"][s, k]
        S = orc.run_sp(m.records(), int(g["N"]), int(g["M"]), g["f"], method="lstsq")
        assert np.max(np.abs(S - g["mc_S"][s])) < 5e-14


def test_known_answers():
    g = golden("ka_divider")
    S = g["S"][:, :, 0]
    assert abs(S[0, 0] - 0.25925925925925924) < 1e-12
    assert abs(S[1, 1] + 0.1111111111111111) < 1e-12
    assert abs(S[1, 0] - 0.9072184232530289) < 1e-12 and abs(S[0, 1] - S[1, 0]) < 1e-13
    g = golden("ka_line")
    assert np.allclose(np.abs(g["S"][1, 0]), 1.0, atol=1e-12)  # matched lossless line
    assert np.allclose(np.abs(g["S"][0, 0]), 0.0, atol=1e-12)


def test_live_reference_check_script_runs_when_the_reference_is_here():
    """`oracle/check_oracle_vs_ref.py` (the open-ended version of the golden pinning) on a few fresh seeds; on a
    machine without the reference tree it reports that and exits 0."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, NUMBA_CACHE_DIR=os.environ.get("NUMBA_CACHE_DIR", "/tmp/numba_cache_heavi_ref"))
    out = subprocess.run([sys.executable, os.path.join(root, "oracle", "check_oracle_vs_ref.py"), "--seeds", "200:203"],
                         capture_output=True, text=True, env=env, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert "all within the gate" in out.stdout or "nothing to check" in out.stdout
