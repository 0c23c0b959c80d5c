"""GPU (-m gpu): several devices in one process (`hvb_run_host_multi`) and the moment
merge of sharded Monte-Carlo statistics.  The splitting logic is exercised on ONE GPU too (two plan replicas
on the same device); with >= 2 visible GPUs the public API is checked end to end."""
import os

import numpy as np
import pytest

import circuits
import heavi_b200 as hv
from heavi_b200 import engine
from heavi_b200.engine import CompiledNetwork, DevicePlanHandle
from oracle import mna_oracle as orc

pytestmark = pytest.mark.gpu


def _ndev():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("case", ["freq", "sample", "table"])
def test_replicas_split_the_job_and_fill_one_array_bit_identically(case):
    if case == "sample":
        m, f, mc = circuits.build_c4(hv, nF=257)
        np.random.seed(2)
        drivers, values = mc._random_numbers, mc.draw(7)
    elif case == "table":
        m = hv.Model(suppress_loadbar=True)
        a, b = m.quick_port(50), m.quick_port(50)
        mid = m.node()
        m.impedance(a, mid, lambda fr: 4.0 + 1e-9j * fr)          # host-tabulated: every slab needs its own columns
        m.inductor(mid, b, 2e-9)
        m.capacitor(mid, m.gnd, 1e-12)
        f, drivers, values = np.linspace(1e9, 3e9, 5003), None, None
    else:
        m, f = circuits.build_c2(hv, "bandpass", nF=70_001)
        drivers, values = None, None
    cn = m._compiled()
    plan, handle = cn.variant(10 ** 9)
    consts, prefs, tables, samples = cn.resolve(f, drivers, values, plan)
    one, st1 = handle.run_host(f, consts, prefs, tables, samples)
    dev2 = 1 if _ndev() > 1 else 0
    reps = [DevicePlanHandle(plan, dev2, kernel_hint=handle.kernel_hint), DevicePlanHandle(plan, 0, kernel_hint=handle.kernel_hint)]
    f32 = engine.pinned_empty(f.shape, np.float32)
    multi, st2 = handle.run_host(f, consts, prefs, tables, samples, f32_out=f32, replicas=reps)
    assert st1 == 0 and st2 == 0
    assert np.array_equal(one, multi)
    assert np.array_equal(f32, f.astype(np.float32))


def test_slab_written_with_a_pitch_into_a_larger_array():
    m, f = circuits.build_c5(hv, nF=3000)
    cn = m._compiled()
    whole = cn.run(f).copy()
    big = engine.pinned_empty((1, 8, 8, 3000))
    big[:] = 0
    lo, hi = 700, 2300
    cn.run(f[lo:hi], out_ptr=big.__array_interface__["data"][0] + lo * 16, out_ld=3000, batch_points=3000)
    assert np.array_equal(big[..., lo:hi], whole[..., lo:hi])
    assert not big[..., :lo].any() and not big[..., hi:].any()


@pytest.mark.skipif("_ndev() < 2")
def test_public_api_uses_all_visible_devices():
    old = os.environ.get("HVB_DEVICES")
    try:
        os.environ["HVB_DEVICES"] = "1"
        m1, f = circuits.build_c5(hv, nF=600_000)
        S1 = m1.run_sp(f)._S.copy()
        os.environ["HVB_DEVICES"] = "all"
        m2, _ = circuits.build_c5(hv, nF=8)
        S2 = m2.run_sp(f)._S
        assert len(m2._compiled().__dict__.get("_replica_cache", {})) >= 1
        assert np.array_equal(S1, S2)
        m3, f3, mc = circuits.build_c4(hv, nF=501)
        np.random.seed(4)
        vals = mc.draw(800)
        mom = m3._compiled().run_stats(f3, mc._random_numbers, vals)
        os.environ["HVB_DEVICES"] = "1"
        m4, _, mc4 = circuits.build_c4(hv, nF=501)
        mom1 = m4._compiled().run_stats(f3, mc4._random_numbers, vals)
        assert np.max(np.abs(mom - mom1)) < 1e-12
        # against the oracle's per-sample S (numeric.py:498-507, sparam.py:215-243)
        sub = slice(0, None, 50)
        Ss = []
        for row in vals[:40]:
            for d, v in zip(mc4._random_numbers, row):
                d._value = float(v)
            Ss.append(orc.run_sp(m4.records(), m4.n_nodes, len(m4.sources), f3[sub], method="lu_batched"))
        os.environ["HVB_DEVICES"] = "all"
        mom40 = m3._compiled().run_stats(f3[sub], mc._random_numbers, vals[:40])
        assert np.max(np.abs(mom40[0] - np.abs(np.stack(Ss)).mean(axis=0))) < 1e-12
    finally:
        if old is None:
            os.environ.pop("HVB_DEVICES", None)
        else:
            os.environ["HVB_DEVICES"] = old
