"""CPU, build container only: the binding stub of INTEGRATION.md, executed verbatim against the UNMODIFIED
reference (`oracle/ref_shim.py`).  A netlist built with reference objects must freeze into the same device
program as the same netlist built with the product front-end -- in particular reference `Scalar` / `Random` /
line-phase closures / Touchstone `interp1d` traces must not fall back to host-tabulated tables."""
import os
import re

import numpy as np
import pytest

import circuits
import heavi_b200 as hv
from heavi_b200 import plan as pl
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not present (GPU box)")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def stub():
    ref = ref_shim.load_reference()
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    code = re.search(r"```python\n(.*?)```", text, flags=re.S).group(1)
    from heavi.base.component import SimulationType          # names network.py already has in scope
    ns = {"SimulationType": SimulationType, "Sparameters": ref.Sparameters}
    exec(compile(code, "INTEGRATION.md", "exec"), ns)
    return ref, ns


def _freeze_reference(ref, ns, build):
    from heavi.base.component import SimulationType
    out = build(ref)
    m = out[0]
    m._assign_internal_nodes(SimulationType.SP)
    m._define_indices(SimulationType.SP)
    m._initialized = True
    return out, ns["hvb_stamp_table"](m)


def _same_program(a: pl.DevicePlan, b: pl.DevicePlan):
    assert (a.mode, a.n, a.ncols, a.P, a.n_real, a.band) == (b.mode, b.n, b.ncols, b.P, b.n_real, b.band)
    for name in ("ops", "coops", "cell_ptr", "cell_ent", "row_ptr", "row_ent", "port_rows", "port_zpref", "fop_code",
                 "fop_out", "prefs", "spl_trace"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
    for name in ("fop_arg", "const_vals", "consts", "epi_scale", "spl_t", "spl_c", "spl_fill", "spl_range"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
    assert a.n_tables == b.n_tables and a.n_samples_cols == b.n_samples_cols and a.epi_simple == b.epi_simple


@pytest.mark.parametrize("name,build,order", [
    ("c2_bandpass", lambda h: circuits.build_c2(h, "bandpass", 8), "rcm"),
    ("c3", lambda h: circuits.build_c3(h, nF=8), "auto"),
    ("c5", lambda h: circuits.build_c5(h, nF=8), "auto"),
])
def test_reference_objects_freeze_into_the_product_plan(stub, name, build, order):
    ref, ns = stub
    _, table_ref = _freeze_reference(ref, ns, build)
    m = build(hv)[0]
    table = m.stamp_table()
    assert table_ref.kinds == table.kinds and table_ref.ids == table.ids and np.array_equal(table_ref.port_table, table.port_table)
    a = pl.compile_plan(table_ref, "auto", order=order)
    b = pl.compile_plan(table, "auto", order=order)
    assert a.n_tables == 0                                   # nothing is tabulated on the host
    _same_program(a, b)


def test_readme_demo_and_monte_carlo_through_reference_objects(stub):
    """c1: the reference's SMD part is a plain Python callable -> exactly one table (the product's own SMD class
    is a closed form); everything else identical.  c4: reference `Random`s become sample columns."""
    ref, ns = stub
    _, t1 = _freeze_reference(ref, ns, lambda h: circuits.build_c1(h, nF=8))
    a = pl.compile_plan(t1, "auto", order="rcm")
    b = pl.compile_plan(circuits.build_c1(hv, nF=8)[0].stamp_table(), "auto", order="rcm")
    assert a.n_tables == 1 and b.n_tables == 0
    assert np.array_equal(a.row_ent, b.row_ent) and np.array_equal(a.fop_code, b.fop_code) and np.array_equal(a.fop_arg, b.fop_arg)
    (m4, f4, mc4), t4 = _freeze_reference(ref, ns, lambda h: circuits.build_c4(h, nF=8))
    a4 = pl.compile_plan(t4, "auto", order="rcm")
    b4 = pl.compile_plan(circuits.build_c4(hv, nF=8)[0].stamp_table(), "auto", order="rcm")
    assert a4.n_tables == 0 and a4.n_samples_cols == 10 == b4.n_samples_cols
    assert all(o is r for o, r in zip(a4.sample_owners, mc4._random_numbers))
    assert np.array_equal(a4.fop_code, b4.fop_code) and np.array_equal(a4.row_ent, b4.row_ent)


def test_wrapped_and_optimiser_parameters(stub):
    ref, ns = stub
    from heavi.lib.optim import OptimVar
    from heavi.numeric import Param
    m = ref.Model(suppress_loadbar=True)
    a, b = m.quick_port(50), m.quick_port(50)
    mid = m.node()
    v, p = OptimVar(2e-9), Param(np.linspace(1e-12, 2e-12, 3))
    p.initialize()                                            # the reference evaluates a value at construction
    m.inductor(a, mid, v)
    m.capacitor(mid, m.gnd, p)
    m.admittance(mid, b, p.inverse())
    m.admittance(mid, m.gnd, p.negative())
    _, t = _freeze_reference(ref, ns, lambda h: (m,))
    plan = pl.compile_plan(t, "auto")
    kinds = sorted(set(int(k) for k in plan.prefs[:, 0]))
    assert plan.n_tables == 0 and plan.n_samples_cols == 2 and pl.PK_SAMPLE_INV in kinds and pl.PK_SAMPLE_NEG in kinds
    assert plan.sample_owners[0] is v and plan.sample_owners[1] is p
