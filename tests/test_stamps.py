"""CPU: the stamp-table compiler below the C-ABI (csrc/hvb_stamps.cpp, `hvb_stamps_compile_host`) against the Python
compiler (heavi_b200/plan.py:compile_plan) -- every array of the device plan, entry by entry, on every golden
circuit, solver mode and padding shape."""
import numpy as np
import pytest

import circuits
import heavi_b200 as hv
from heavi_b200 import engine
from heavi_b200 import plan as pl
from helpers import registry

REG = registry(hv)
FIELDS_EXACT = ("ops", "coops", "cell_ptr", "cell_ent", "row_ptr", "row_ent", "fop_code", "fop_out", "fop_arg", "port_rows", "port_zpref",
                "prefs", "consts", "const_vals", "epi_port_of_row")


def same(py: pl.DevicePlan, c: dict, band=True):
    assert (py.mode, py.n, py.ncols, py.P, py.n_real, py.n_dyn, py.coop_scratch, py.epi_simple, py.max_nport) == \
           (c["mode"], c["n"], c["ncols"], c["P"], c["n_real"], c["n_dyn"], c["coop_scratch"], c["epi_simple"], c["max_nport"])
    assert py.n_tables == c["n_tables"] and py.n_samples_cols == c["n_samples_cols"]
    for name in FIELDS_EXACT:
        a, b = np.asarray(getattr(py, name)), np.asarray(c[name])
        assert a.shape == b.shape and np.array_equal(a, b, equal_nan=True), name
    assert np.array_equal(np.asarray(py.epi_scale).reshape(-1), c["epi_scale"])
    assert [t[0] for t in py.table_params] == list(c["table_pref"]) and [t[1] for t in py.table_params] == list(c["table_slot"])
    if band:
        assert py.band == c["band"]


@pytest.mark.parametrize("name", sorted(REG))
def test_library_compiler_equals_python_compiler(name):
    m = REG[name]()
    table = m.stamp_table()
    tagged = pl.tag_table(table)
    modes = ["full"] + (["norton"] if pl.norton_legal(table) else [])
    for mode in modes:
        py = pl.compile_plan(table, mode, order="natural")
        same(py, engine.stamps_compile_host(tagged, mode, order="natural"))
        shape = (py.n_real + 3, py.P + 2)
        same(pl.compile_plan(table, mode, pad_to=shape), engine.stamps_compile_host(tagged, mode, pad_to=shape))
    same(pl.compile_plan(table, "auto", order="natural"), engine.stamps_compile_host(tagged, "auto", order="natural"))
    same(pl.compile_plan(table, "dump"), engine.stamps_compile_host(tagged, "dump"))


def test_sample_table_and_closed_form_parameters():
    m, f, mc = circuits.build_c4(hv, nF=8)
    same(pl.compile_plan(m.stamp_table(), "auto", order="natural"), engine.stamps_compile_host(pl.tag_table(m.stamp_table()), "auto", order="natural"))
    m2 = hv.Model(suppress_loadbar=True)
    a, b = m2.quick_port(50), m2.quick_port(75)
    mid, mid2 = m2.node(), m2.node()
    m2.impedance(a, mid, lambda fr: 3.0 + 1j * 2e-9 * fr)
    m2.transmissionline(mid, mid2, 60.0, 3.1 - 0.05j, 0.011)
    m2.admittance(mid2, b, lambda fr: 0.02 + 1j * 1e-12 * fr)
    sweep = hv.ParameterSweep()
    Cp = sweep.lin(1e-12, 2e-12).add(3)
    Cp.initialize()
    m2.capacitor(mid, m2.gnd, Cp)
    m2.admittance(mid, m2.gnd, Cp.inverse())
    same(pl.compile_plan(m2.stamp_table(), "auto", order="natural"), engine.stamps_compile_host(pl.tag_table(m2.stamp_table()), "auto", order="natural"))
    m3, _ = circuits.build_random_twoport(hv)[:2]
    same(pl.compile_plan(m3.stamp_table(), "auto", order="natural"), engine.stamps_compile_host(pl.tag_table(m3.stamp_table()), "auto", order="natural"))


@pytest.mark.parametrize("build", [lambda: circuits.build_c5(hv, nF=8)[0], lambda: circuits.build_c2(hv, "bandpass", 8)[0],
                                   lambda: circuits.build_band_ladder(hv, K=60, reach=3, nports=3)[0]])
def test_bandwidth_reducing_order_is_a_valid_narrow_permutation(build):
    m = build()
    table = m.stamp_table()
    py = pl.compile_plan(table, "auto", order="rcm")
    c = engine.stamps_compile_host(pl.tag_table(table), "auto", order="rcm")
    assert c["band"] >= 1 and c["band"] <= max(py.band, 1) + 1              # own Cuthill-McKee: as narrow as scipy's (+-1)
    assert c["n"] == py.n and sorted(c["port_rows"][:, 0]) != [] and len(set(c["port_rows"][:, 0])) == py.P
    # the matrix really is banded with that half width in the library's order
    for r in range(c["n"]):
        for e in range(c["row_ptr"][r], c["row_ptr"][r + 1]):
            col = (int(c["row_ent"][e]) >> 20) & 0x7FF
            assert col >= c["n"] or abs(col - r) <= c["band"]
