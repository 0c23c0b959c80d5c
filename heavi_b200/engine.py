"""ctypes binding of `libheavi_b200.so` (C-ABI: include/heavi_b200.h) and the compiled-netlist object
`Network.run_sp` drives.

There is deliberately no CPU path here: if the CUDA library is missing or no CUDA device is
present, every entry raises.  PyTorch is used only for pinned host buffers and (in the
device-resident entry used by the benchmark) for device buffers and streams; no tensor crosses the
C-ABI, only `data_ptr()` integers.
"""
from __future__ import annotations

import ctypes as C
import os
import warnings

import numpy as np

from . import plan as pl

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libheavi_b200.so")

HVB_STATUS_NONFINITE = 1
HVB_STATUS_ZERO_PIVOT = 2
HVB_STATUS_RERUN = 4

_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)


class PlanDesc(C.Structure):
    _fields_ = [(k, C.c_int32) for k in (
        "n", "ncols", "P", "n_consts", "n_prefs", "n_const_vals", "n_dyn", "n_ops", "n_coops",
        "n_samples_cols", "coop_scratch", "n_traces", "n_knots", "n_coefs", "n_row_ent", "n_fast_ops", "epi_simple", "kernel_hint")] + [
        ("const_vals", _f64p), ("ops", _i32p), ("coops", _i32p), ("cell_ptr", _i32p), ("cell_ent", _i32p),
        ("row_ptr", _i32p), ("row_ent", _i32p), ("fop_code", _i32p), ("fop_out", _i32p), ("fop_arg", _f64p),
        ("epi_port_of_row", _i32p), ("epi_scale", _f64p), ("port_rows", _i32p), ("port_zpref", _i32p), ("spl_t", _f64p), ("spl_c", _f64p),
        ("spl_trace", _i32p), ("spl_fill", _f64p), ("spl_range", _f64p)]


class RunArgs(C.Structure):
    _fields_ = [("consts", C.c_void_p), ("prefs", C.c_void_p), ("f", C.c_void_p), ("nF", C.c_int64),
                ("samples", C.c_void_p), ("nS", C.c_int64), ("tables", C.c_void_p), ("n_tables", C.c_int32),
                ("f32_out", C.c_void_p), ("out_ld", C.c_int64)]


class KernelInfo(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("kernel_class", "group", "pmax", "threads", "smem_bytes", "regs",
                                         "ctas_per_sm", "sm_count", "rows_per_lane", "stream_words")] + [("launches", C.c_int64)] + [
        (k, C.c_int32) for k in ("gen_rolled", "gen_cases", "local_bytes", "gen_scratch_acc")] + [("gen_flops_per_solve", C.c_int64)] + [
        ("gen_hub", C.c_int32), ("gen_two_sided", C.c_int32)]


class CodegenInfo(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("supported", "rolled", "n_cases", "scratch_acc")] + [
        (k, C.c_int64) for k in ("src_len", "n_rows", "n_itab", "n_dtab", "flops_per_solve")] + [
        ("hub", C.c_int32), ("threads", C.c_int32), ("smem_bytes", C.c_int64), ("two_sided", C.c_int32), ("cut_row", C.c_int32),
        ("reason", C.c_char * 256)]


EXPORTS = ("hvb_warp_shape", "hvb_plan_create", "hvb_plan_destroy", "hvb_run_host", "hvb_run_device",
           "hvb_run_stats_host", "hvb_assemble_host",
           "hvb_plan_kernel_info", "hvb_measure_dfma_peak", "hvb_last_error", "hvb_version",
           "hvb_codegen_source", "hvb_codegen_compile", "hvb_plan_set_tunable", "hvb_run_host_multi",
           "hvb_run_host_async", "hvb_run_host_wait", "hvb_plan_create_from_stamps", "hvb_stamps_compile_host",
           "hvb_compiled_free")

_lib = None


def load_library():
    """dlopen the CUDA library; raises (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C heavi_b200/csrc`.  heavi_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.hvb_plan_create.argtypes = [C.POINTER(PlanDesc), C.c_int, C.POINTER(C.c_void_p)]
    lib.hvb_plan_create.restype = C.c_int
    lib.hvb_warp_shape.argtypes = [C.c_int32, C.c_int32, C.c_int32, _i32p, _i32p]
    lib.hvb_warp_shape.restype = C.c_int
    lib.hvb_plan_destroy.argtypes = [C.c_void_p]
    lib.hvb_plan_destroy.restype = None
    lib.hvb_run_host.argtypes = [C.c_void_p, C.POINTER(RunArgs), C.c_void_p, _i32p]
    lib.hvb_run_host.restype = C.c_int
    lib.hvb_plan_create_from_stamps.argtypes = [C.c_void_p, C.c_int32, C.c_int, C.POINTER(C.c_void_p)]
    lib.hvb_plan_create_from_stamps.restype = C.c_int
    lib.hvb_stamps_compile_host.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.hvb_stamps_compile_host.restype = C.c_int
    lib.hvb_compiled_free.argtypes = [C.c_void_p]
    lib.hvb_compiled_free.restype = None
    lib.hvb_run_host_async.argtypes = [C.c_void_p, C.POINTER(RunArgs), C.c_void_p]
    lib.hvb_run_host_async.restype = C.c_int
    lib.hvb_run_host_wait.argtypes = [C.c_void_p, _i32p]
    lib.hvb_run_host_wait.restype = C.c_int
    lib.hvb_run_host_multi.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.POINTER(RunArgs), C.c_void_p, _i32p]
    lib.hvb_run_host_multi.restype = C.c_int
    lib.hvb_run_stats_host.argtypes = [C.c_void_p, C.POINTER(RunArgs), C.c_void_p, _i32p]
    lib.hvb_run_stats_host.restype = C.c_int
    lib.hvb_run_device.argtypes = [C.c_void_p, C.POINTER(RunArgs), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    lib.hvb_run_device.restype = C.c_int
    lib.hvb_assemble_host.argtypes = [C.c_void_p, C.POINTER(RunArgs), C.c_void_p]
    lib.hvb_assemble_host.restype = C.c_int
    lib.hvb_plan_kernel_info.argtypes = [C.c_void_p, C.POINTER(KernelInfo)]
    lib.hvb_plan_kernel_info.restype = C.c_int
    lib.hvb_measure_dfma_peak.argtypes = [C.c_int, _f64p]
    lib.hvb_measure_dfma_peak.restype = C.c_int
    lib.hvb_codegen_source.argtypes = [C.POINTER(PlanDesc), C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.POINTER(CodegenInfo)]
    lib.hvb_codegen_source.restype = C.c_int
    lib.hvb_codegen_compile.argtypes = [C.c_char_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    lib.hvb_codegen_compile.restype = C.c_int
    lib.hvb_plan_set_tunable.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
    lib.hvb_plan_set_tunable.restype = C.c_int
    lib.hvb_last_error.restype = C.c_char_p
    lib.hvb_version.restype = C.c_char_p
    _lib = lib
    return lib


def _check(rc: int):
    if rc != 0:
        msg = load_library().hvb_last_error().decode()
        if rc == -1:
            raise ValueError(f"heavi_b200: {msg}")
        if rc == -3:
            raise NotImplementedError(f"heavi_b200: {msg}")
        raise RuntimeError(f"heavi_b200 (code {rc}): {msg}")


def current_device() -> int:
    env = os.environ.get("HVB_DEVICE")
    if env is not None:
        return int(env)
    try:
        import torch
        if torch.cuda.is_available():
            return torch.cuda.current_device()
    except Exception:
        pass
    if "LOCAL_RANK" in os.environ:
        return int(os.environ["LOCAL_RANK"])
    return 0


def visible_devices() -> list[int]:
    """CUDA ordinals one `run_sp` call may use.  One process per GPU (torchrun: WORLD_SIZE > 1) or an explicit
    `HVB_DEVICE` pins the process to its own device; otherwise `HVB_DEVICES` ("all", a count, or a comma list)
    decides, default "all" -- a plain `Model.run_sp(f)` on an 8-GPU box uses the 8 GPUs for large sweeps."""
    first = current_device()
    if os.environ.get("HVB_DEVICE") is not None or int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return [first]
    spec = os.environ.get("HVB_DEVICES", "all").strip().lower()
    try:
        import torch
        count = torch.cuda.device_count()
    except Exception:
        count = 1
    if spec in ("", "all"):
        devs = list(range(count))
    elif "," in spec:
        devs = [int(x) for x in spec.split(",") if x.strip() != ""]
    else:
        devs = list(range(min(int(spec), count)))
    devs = [d for d in devs if 0 <= d < max(count, 1)] or [first]
    if first in devs:                                   # the caller's current device leads (it owns block 0)
        devs.remove(first)
    return [first] + devs


def pinned_empty(shape, dtype=np.complex128) -> np.ndarray:
    """numpy array on page-locked memory (torch's caching host allocator), so that the D2H of S runs
    at full PCIe rate and asynchronously."""
    import torch
    tdt = {np.dtype(np.complex128): torch.complex128, np.dtype(np.float64): torch.float64,
           np.dtype(np.float32): torch.float32}[np.dtype(dtype)]
    return torch.empty(tuple(int(x) for x in shape), dtype=tdt, pin_memory=True).numpy()


def _ptr(a: np.ndarray):
    """Address of a numpy array as a plain integer (what a c_void_p field takes); the array interface is
    several times cheaper than building a ctypes pointer object on every call."""
    return a.__array_interface__["data"][0] if a is not None and a.size else None


def make_desc(plan: pl.DevicePlan, kernel_hint: int = 0):
    """(PlanDesc, arrays it points into) for a compiled plan."""
    keep = {}

    def arr(name, a, dt):
        a = np.ascontiguousarray(a, dtype=dt)
        if a.size == 0:
            a = np.zeros(1, dtype=dt)
        keep[name] = a
        return a

    if True:
        d = PlanDesc()
        d.n, d.ncols, d.P = plan.n, plan.ncols, plan.P
        d.n_consts, d.n_prefs = len(plan.consts), len(plan.prefs)
        d.n_const_vals, d.n_dyn = len(plan.const_vals), plan.n_dyn
        d.n_ops, d.n_coops = len(plan.ops), len(plan.coops)
        d.n_samples_cols, d.coop_scratch = plan.n_samples_cols, plan.coop_scratch
        d.n_traces, d.n_knots, d.n_coefs = len(plan.spl_trace), len(plan.spl_t), len(plan.spl_c)
        d.const_vals = arr("cv", plan.const_vals, np.complex128).ctypes.data_as(_f64p)
        d.ops = arr("ops", plan.ops, np.int32).ctypes.data_as(_i32p)
        d.coops = arr("coops", plan.coops, np.int32).ctypes.data_as(_i32p)
        d.cell_ptr = arr("cp", plan.cell_ptr, np.int32).ctypes.data_as(_i32p)
        d.cell_ent = arr("ce", plan.cell_ent, np.int32).ctypes.data_as(_i32p)
        d.n_row_ent = len(plan.row_ent)
        d.row_ptr = arr("rp", plan.row_ptr, np.int32).ctypes.data_as(_i32p)
        d.row_ent = arr("re", plan.row_ent, np.int32).ctypes.data_as(_i32p)
        d.n_fast_ops, d.epi_simple = len(plan.fop_code), plan.epi_simple
        d.kernel_hint = int(kernel_hint)
        d.fop_code = arr("fc", plan.fop_code, np.int32).ctypes.data_as(_i32p)
        d.fop_out = arr("fo", plan.fop_out, np.int32).ctypes.data_as(_i32p)
        d.fop_arg = arr("fa", plan.fop_arg, np.float64).ctypes.data_as(_f64p)
        d.epi_port_of_row = arr("epr", plan.epi_port_of_row, np.int32).ctypes.data_as(_i32p)
        d.epi_scale = arr("eps", plan.epi_scale, np.float64).ctypes.data_as(_f64p)
        d.port_rows = arr("pr", plan.port_rows, np.int32).ctypes.data_as(_i32p)
        d.port_zpref = arr("pz", plan.port_zpref, np.int32).ctypes.data_as(_i32p)
        d.spl_t = arr("st", plan.spl_t, np.float64).ctypes.data_as(_f64p)
        d.spl_c = arr("sc", plan.spl_c, np.complex128).ctypes.data_as(_f64p)
        d.spl_trace = arr("str", plan.spl_trace, np.int32).ctypes.data_as(_i32p)
        d.spl_fill = arr("sf", plan.spl_fill, np.complex128).ctypes.data_as(_f64p)
        d.spl_range = arr("sr", plan.spl_range, np.float64).ctypes.data_as(_f64p)
    return d, keep


def codegen_source(plan: pl.DevicePlan, prefs: np.ndarray | None = None, mode: int = 0, consts: np.ndarray | None = None,
                   bake_consts: bool = True) -> dict:
    """Source of the kernel the library would generate for `plan` (kernel class 3), without touching a
    GPU: {"supported", "reason", "src", "rolled", "rows", "itab", "dtab", "n_cases", "flops_per_solve"}."""
    lib = load_library()
    d, keep = make_desc(plan, 2)
    prefs = np.ascontiguousarray(plan.prefs if prefs is None else prefs, dtype=np.int32)
    consts = np.ascontiguousarray(plan.consts if consts is None else consts, dtype=np.complex128)
    cptr = _ptr(consts) if bake_consts else None      # what the library does at run time: constants baked into the code
    info = CodegenInfo()
    _check(lib.hvb_codegen_source(C.byref(d), _ptr(prefs), cptr, mode, None, 0, None, None, None, C.byref(info)))
    out = {"supported": bool(info.supported), "reason": info.reason.decode(), "rolled": bool(info.rolled),
           "n_cases": info.n_cases, "scratch_acc": bool(info.scratch_acc), "flops_per_solve": info.flops_per_solve,
           "hub": bool(info.hub), "threads": info.threads, "smem_bytes": info.smem_bytes,
           "two_sided": bool(info.two_sided), "cut_row": info.cut_row}
    if not info.supported:
        return out
    src = C.create_string_buffer(info.src_len + 1)
    rows = np.zeros(max(info.n_rows, 1), dtype=np.int32)
    itab = np.zeros(max(info.n_itab, 1), dtype=np.int32)
    dtab = np.zeros(max(info.n_dtab, 1), dtype=np.float64)
    _check(lib.hvb_codegen_source(C.byref(d), _ptr(prefs), cptr, mode, C.cast(src, C.c_void_p), info.src_len + 1,
                                  _ptr(rows), _ptr(itab), _ptr(dtab), C.byref(info)))
    out.update(src=src.value.decode(), rows=rows[:info.n_rows], itab=itab[:info.n_itab], dtab=dtab[:info.n_dtab],
               flops_per_solve=info.flops_per_solve, hub=bool(info.hub), threads=info.threads, smem_bytes=info.smem_bytes,
               two_sided=bool(info.two_sided), cut_row=info.cut_row, n_cases=info.n_cases)
    return out


class ParamC(C.Structure):
    _fields_ = [("kind", C.c_int32), ("index", C.c_int32), ("v", C.c_double * 4)]


class StampTableC(C.Structure):
    _fields_ = [("n_components", C.c_int32), ("comp_kind", _i32p), ("comp_port", _i32p), ("ids_off", _i32p), ("ids", _i32p),
                ("par_off", _i32p), ("par", C.POINTER(ParamC)), ("P", C.c_int32), ("port_table", _i32p), ("port_src", _i32p),
                ("N", C.c_int32), ("M", C.c_int32), ("n_traces", C.c_int32), ("n_knots", C.c_int32), ("n_coefs", C.c_int32),
                ("spl_t", _f64p), ("spl_c", _f64p), ("spl_trace", _i32p), ("spl_fill", _f64p), ("spl_range", _f64p)]


MODES = {"auto": 0, "full": 1, "norton": 2, "dump": 3}
ORDERS = {"auto": 0, "natural": 1, "rcm": 2}


def stamp_table_c(tagged: "pl.TaggedTable"):
    """(StampTableC, keep-alive) for a `plan.TaggedTable`."""
    keep = {}

    def arr(name, a, dt):
        a = np.ascontiguousarray(a, dtype=dt)
        if a.size == 0:
            a = np.zeros(1, dtype=dt)
        keep[name] = a
        return a

    t = StampTableC()
    t.n_components = len(tagged.comp_kind)
    t.comp_kind = arr("ck", tagged.comp_kind, np.int32).ctypes.data_as(_i32p)
    t.comp_port = arr("cp", tagged.comp_port, np.int32).ctypes.data_as(_i32p)
    t.ids_off = arr("io", tagged.ids_off, np.int32).ctypes.data_as(_i32p)
    t.ids = arr("ids", tagged.ids, np.int32).ctypes.data_as(_i32p)
    t.par_off = arr("po", tagged.par_off, np.int32).ctypes.data_as(_i32p)
    par = np.ascontiguousarray(tagged.par)
    assert par.dtype.itemsize == C.sizeof(ParamC)
    keep["par"] = par
    t.par = C.cast(par.ctypes.data, C.POINTER(ParamC))
    t.P = tagged.port_table.shape[0]
    t.port_table = arr("pt", tagged.port_table, np.int32).ctypes.data_as(_i32p)
    t.port_src = arr("ps", tagged.port_src, np.int32).ctypes.data_as(_i32p)
    t.N, t.M = tagged.N, tagged.M
    t.n_traces, t.n_knots, t.n_coefs = len(tagged.spl_trace), len(tagged.spl_t), len(tagged.spl_c)
    t.spl_t = arr("st", tagged.spl_t, np.float64).ctypes.data_as(_f64p)
    t.spl_c = arr("sc", tagged.spl_c, np.complex128).ctypes.data_as(_f64p)
    t.spl_trace = arr("str", tagged.spl_trace, np.int32).ctypes.data_as(_i32p)
    t.spl_fill = arr("sf", tagged.spl_fill, np.complex128).ctypes.data_as(_f64p)
    t.spl_range = arr("sr", tagged.spl_range, np.float64).ctypes.data_as(_f64p)
    return t, keep


def stamps_compile_host(tagged: "pl.TaggedTable", mode: str = "auto", pad_to=None, order: str = "auto") -> dict:
    """The plan the library's own compiler (csrc/hvb_stamps.cpp) produces for a tagged stamp table, as numpy arrays
    named like `plan.DevicePlan`'s fields (no CUDA call)."""
    lib = load_library()
    t, keep = stamp_table_c(tagged)
    d = PlanDesc()
    handle, consts, prefs = C.c_void_p(), _f64p(), _i32p()
    tp, ts = _i32p(), _i32p()
    info = (C.c_int32 * 8)()
    _check(lib.hvb_stamps_compile_host(C.byref(t), MODES[mode], pad_to[0] if pad_to else 0, pad_to[1] if pad_to else 0, ORDERS[order],
                                       C.byref(handle), C.byref(d), C.byref(consts), C.byref(prefs), info, C.byref(tp), C.byref(ts)))
    try:
        def take(ptr, n, dt):
            return np.ctypeslib.as_array(ptr, shape=(max(n, 1),)).copy()[:n].astype(dt) if n else np.zeros(0, dtype=dt)
        n, ncols = d.n, d.ncols
        nnz = int(np.ctypeslib.as_array(d.cell_ptr, shape=(n * ncols + 1,))[-1])
        out = dict(
            mode={1: "full", 2: "norton", 3: "dump"}[info[0]], n=n, ncols=ncols, P=d.P, n_real=info[1], band=info[2], max_nport=info[3],
            n_tables=info[4], n_samples_cols=info[5], n_dyn=d.n_dyn, coop_scratch=d.coop_scratch, epi_simple=d.epi_simple,
            consts=take(consts, d.n_consts * 2, np.float64).view(np.complex128) if d.n_consts else np.zeros(0, dtype=np.complex128),
            prefs=take(prefs, d.n_prefs * 2, np.int32).reshape(-1, 2),
            const_vals=take(d.const_vals, d.n_const_vals * 2, np.float64).view(np.complex128),
            ops=take(d.ops, d.n_ops * 8, np.int32).reshape(-1, 8), coops=take(d.coops, d.n_coops * 8, np.int32).reshape(-1, 8),
            cell_ptr=take(d.cell_ptr, n * ncols + 1, np.int32), cell_ent=take(d.cell_ent, nnz, np.int32),
            row_ptr=take(d.row_ptr, n + 1, np.int32), row_ent=take(d.row_ent, d.n_row_ent, np.int32),
            fop_code=take(d.fop_code, d.n_fast_ops, np.int32), fop_out=take(d.fop_out, d.n_fast_ops, np.int32),
            fop_arg=take(d.fop_arg, d.n_fast_ops, np.float64),
            epi_port_of_row=take(d.epi_port_of_row, max(n, 1), np.int32), epi_scale=take(d.epi_scale, max(d.P, 1) ** 2, np.float64),
            port_rows=take(d.port_rows, d.P * 3, np.int32).reshape(-1, 3), port_zpref=take(d.port_zpref, d.P, np.int32),
            table_pref=take(tp, info[6], np.int32), table_slot=take(ts, info[6], np.int32))
    finally:
        lib.hvb_compiled_free(handle)
    return out


class StampPlanHandle:
    """A plan created by the library's own compiler from a tagged stamp table (`hvb_plan_create_from_stamps`): the
    path a host without Python takes.  Per-run work left to this wrapper: tabulating callables and drawing the
    sample matrix (`plan.resolve_run` on a light stand-in for `plan.DevicePlan`)."""

    def __init__(self, table: "pl.StampTable", mode: str = "auto", device: int | None = None):
        self.lib = load_library()
        self.device = current_device() if device is None else device
        self.tagged = pl.tag_table(table)
        t, self._keep = stamp_table_c(self.tagged)
        handle = C.c_void_p()
        _check(self.lib.hvb_plan_create_from_stamps(C.byref(t), MODES[mode], self.device, C.byref(handle)))
        self.handle = handle
        self.P = t.P

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.hvb_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def kernel_info(self) -> dict:
        ki = KernelInfo()
        _check(self.lib.hvb_plan_kernel_info(self.handle, C.byref(ki)))
        return {k: getattr(ki, k) for k, _ in KernelInfo._fields_}

    def run(self, f, drivers=None, values=None):
        """`run_host` with the per-run host work done: every table callable evaluated on f (a scalar result is
        broadcast), the sample matrix filled from `values[s, k]` = value of `drivers[k]` (or the owners' current
        values)."""
        f = np.ascontiguousarray(f, dtype=np.float64)
        tabs = [np.broadcast_to(np.asarray(fn(f), dtype=np.complex128), f.shape) for fn in self.tagged.table_fns]
        tables = np.ascontiguousarray(np.stack(tabs)) if tabs else None
        owners = self.tagged.sample_owners
        nS = 1 if values is None else len(values)
        samples = np.empty((nS, len(owners)))
        col_of = {} if drivers is None else {id(d): k for k, d in enumerate(drivers)}
        for c, o in enumerate(owners):
            samples[:, c] = np.asarray(values)[:, col_of[id(o)]] if id(o) in col_of else float(o._value)
        return self.run_host(f, tables, samples)

    def run_host(self, f, tables=None, samples=None):
        """S[nS,P,P,nF]; consts / prefs are the plan's own (NULL in the run arguments)."""
        f = np.ascontiguousarray(f, dtype=np.float64)
        nF = f.shape[0]
        samples = np.zeros((1, 0)) if samples is None else np.ascontiguousarray(samples, dtype=np.float64)
        tables = np.zeros((0, nF), dtype=np.complex128) if tables is None else np.ascontiguousarray(tables, dtype=np.complex128)
        nS = samples.shape[0]
        out = pinned_empty((nS, self.P, self.P, nF))
        a = RunArgs()
        a.consts, a.prefs = None, None
        a.f, a.nF = _ptr(f), nF
        a.samples, a.nS = _ptr(samples), nS
        a.tables, a.n_tables = _ptr(tables), tables.shape[0]
        status = C.c_int32(0)
        _check(self.lib.hvb_run_host(self.handle, C.byref(a), C.c_void_p(out.__array_interface__["data"][0]), C.byref(status)))
        return out, status.value


def codegen_compile(src: str) -> str:
    """NVRTC-compile generated source for sm_100a (no GPU needed); returns the compiler log, raises on errors."""
    lib = load_library()
    log = C.create_string_buffer(1 << 16)
    size = C.c_int64(0)
    _check(lib.hvb_codegen_compile(src.encode(), C.cast(log, C.c_void_p), len(log), C.byref(size)))
    return log.value.decode()


class DevicePlanHandle:
    """Owns one `hvb_plan*`."""

    def __init__(self, plan: pl.DevicePlan, device: int | None = None, kernel_hint: int = 0):
        """`kernel_hint`: 0 library decides, 1 dense kernels only, 2 banded / generated kernel when the structure allows."""
        self.lib = load_library()
        self.plan = plan
        self.device = current_device() if device is None else device
        self.kernel_hint = kernel_hint
        d, self._keep = make_desc(plan, kernel_hint)
        handle = C.c_void_p()
        _check(self.lib.hvb_plan_create(C.byref(d), self.device, C.byref(handle)))
        self.handle = handle

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.hvb_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def _args(self, consts, prefs, f_ptr, nF, samples_ptr, nS, tables_ptr, nT):
        self._run_keep = (np.ascontiguousarray(consts, dtype=np.complex128), np.ascontiguousarray(prefs, dtype=np.int32))
        a = RunArgs()
        a.consts = _ptr(self._run_keep[0])
        a.prefs = _ptr(self._run_keep[1])
        a.f, a.nF = f_ptr, nF
        a.samples, a.nS = samples_ptr, nS
        a.tables, a.n_tables = tables_ptr, nT
        return a

    def set_tunable(self, name: str, value: float):
        """Host-pipeline tunable of this plan (`hvb_plan_set_tunable`): chunk_mb, min_slab, stats_mb, ..."""
        _check(self.lib.hvb_plan_set_tunable(self.handle, name.encode(), float(value)))

    def kernel_info(self) -> dict:
        ki = KernelInfo()
        _check(self.lib.hvb_plan_kernel_info(self.handle, C.byref(ki)))
        return {k: getattr(ki, k) for k, _ in KernelInfo._fields_}

    # ---- host buffers in, host buffer out ------------------------------------------------------
    def run_host(self, f, consts, prefs, tables, samples, out: np.ndarray | None = None,
                 f32_out: np.ndarray | None = None, out_ptr: int | None = None, out_ld: int = 0, replicas=None):
        """`f32_out` (float32 [nF], optional) receives the frequencies rounded to float32, filled in by
        the library while the GPU works (the reference's `Sparameters.f`).  `out_ptr` / `out_ld`: write into
        a caller-owned array at that address with frequency pitch `out_ld` (a slab of a larger result).
        `replicas`: handles of the same plan on other devices -> the call is spread over all of them
        (`hvb_run_host_multi`)."""
        p = self.plan
        f = np.ascontiguousarray(f, dtype=np.float64)
        nF, nS = f.shape[0], samples.shape[0]
        if out is None and out_ptr is None:
            out = pinned_empty((nS, p.P, p.P, nF))
        tables = np.ascontiguousarray(tables, dtype=np.complex128)
        samples = np.ascontiguousarray(samples, dtype=np.float64)
        a = self._args(consts, prefs, _ptr(f), nF, _ptr(samples), nS, _ptr(tables), tables.shape[0])
        if f32_out is not None:
            if f32_out.dtype != np.float32 or f32_out.shape != (nF,) or not f32_out.flags.c_contiguous:
                raise ValueError("f32_out must be a contiguous float32 array of nF elements")
            a.f32_out = _ptr(f32_out)
        a.out_ld = int(out_ld)
        dst = C.c_void_p(out_ptr if out_ptr is not None else out.__array_interface__["data"][0])
        status = C.c_int32(0)
        if replicas:
            handles = (C.c_void_p * (1 + len(replicas)))(self.handle, *[r.handle for r in replicas])
            _check(self.lib.hvb_run_host_multi(handles, len(handles), C.byref(a), dst, C.byref(status)))
        else:
            _check(self.lib.hvb_run_host(self.handle, C.byref(a), dst, C.byref(status)))
        return out, status.value

    def run_host_async(self, f, consts, prefs, tables, samples, out: np.ndarray | None = None, f32_out: np.ndarray | None = None):
        """Queue one run on this plan's streams and return at once (`hvb_run_host_async`); `wait()` completes it.
        Returns the output array (valid after `wait`)."""
        p = self.plan
        f = np.ascontiguousarray(f, dtype=np.float64)
        nF, nS = f.shape[0], samples.shape[0]
        if out is None:
            out = pinned_empty((nS, p.P, p.P, nF))
        tables = np.ascontiguousarray(tables, dtype=np.complex128)
        samples = np.ascontiguousarray(samples, dtype=np.float64)
        a = self._args(consts, prefs, _ptr(f), nF, _ptr(samples), nS, _ptr(tables), tables.shape[0])
        if f32_out is not None:
            a.f32_out = _ptr(f32_out)
        self._pending_keep = (f, tables, samples, out, f32_out, a)       # alive until wait()
        _check(self.lib.hvb_run_host_async(self.handle, C.byref(a), C.c_void_p(out.__array_interface__["data"][0])))
        return out

    def wait(self) -> int:
        status = C.c_int32(0)
        _check(self.lib.hvb_run_host_wait(self.handle, C.byref(status)))
        self._pending_keep = None
        return status.value

    def run_stats_host(self, f, consts, prefs, tables, samples):
        """Monte-Carlo moments reduced on the device: float64 [6, P, P, nF]."""
        p = self.plan
        f = np.ascontiguousarray(f, dtype=np.float64)
        nF, nS = f.shape[0], samples.shape[0]
        out = pinned_empty((6, p.P, p.P, nF), dtype=np.float64)
        tables = np.ascontiguousarray(tables, dtype=np.complex128)
        samples = np.ascontiguousarray(samples, dtype=np.float64)
        a = self._args(consts, prefs, _ptr(f), nF, _ptr(samples), nS, _ptr(tables), tables.shape[0])
        status = C.c_int32(0)
        _check(self.lib.hvb_run_stats_host(self.handle, C.byref(a), out.ctypes.data_as(C.c_void_p), C.byref(status)))
        return out, status.value

    def assemble_host(self, f, consts, prefs, tables, samples) -> np.ndarray:
        p = self.plan
        f = np.ascontiguousarray(f, dtype=np.float64)
        nF, nS = f.shape[0], samples.shape[0]
        out = np.empty((nS, nF, p.n, p.ncols), dtype=np.complex128)
        tables = np.ascontiguousarray(tables, dtype=np.complex128)
        samples = np.ascontiguousarray(samples, dtype=np.float64)
        a = self._args(consts, prefs, _ptr(f), nF, _ptr(samples), nS, _ptr(tables), tables.shape[0])
        _check(self.lib.hvb_assemble_host(self.handle, C.byref(a), out.ctypes.data_as(C.c_void_p)))
        return out

    # ---- device buffers (torch tensors) ----------------------------------------------------------
    def run_device(self, d_f, consts, prefs, d_tables, d_samples, d_out, d_status=None, stream=None):
        """All `d_*` are torch CUDA tensors; enqueues on `stream` (torch stream) without syncing."""
        import torch
        nF = d_f.shape[0]
        nS = d_samples.shape[0] if d_samples is not None else 1
        nT = d_tables.shape[0] if d_tables is not None else 0
        st = stream if stream is not None else torch.cuda.current_stream()
        a = self._args(consts, prefs, d_f.data_ptr(), nF,
                       d_samples.data_ptr() if d_samples is not None and d_samples.numel() else None, nS,
                       d_tables.data_ptr() if nT else None, nT)
        _check(self.lib.hvb_run_device(self.handle, C.byref(a), d_out.data_ptr(), d_out.shape[-1],
                                       d_status.data_ptr() if d_status is not None else None,
                                       C.c_void_p(st.cuda_stream)))


def warp_shape(n: int, P: int, max_nport: int = 0):
    """(rows, rhs columns) the register-resident kernel wants this system padded to, or None."""
    lib = load_library()
    a, b = C.c_int32(0), C.c_int32(0)
    _check(lib.hvb_warp_shape(n, P, max_nport, C.byref(a), C.byref(b)))
    return (a.value, b.value) if a.value > 0 else None


def measure_dfma_peak(device: int | None = None) -> float:
    lib = load_library()
    out = C.c_double(0.0)
    _check(lib.hvb_measure_dfma_peak(current_device() if device is None else device, C.byref(out)))
    return out.value


class _DriversSet:
    """Context: batch drivers (Random / Param / optimiser variables) temporarily read the values of one row.
    Drivers whose `_value` is not a plain attribute (OptimVar: derived from its coordinate) offer
    `_set_physical`."""

    def __init__(self, drivers):
        self.drivers = list(drivers)

    def __enter__(self):
        self.saved = [None if hasattr(d, "_set_physical") else d._value for d in self.drivers]
        return self

    def set(self, row):
        for d, v in zip(self.drivers, row):
            if hasattr(d, "_set_physical"):
                d._set_physical(float(v))
            else:
                d._value = float(v)

    def __exit__(self, *exc):
        for d, v in zip(self.drivers, self.saved):
            if hasattr(d, "_set_physical"):
                d._set_physical(None)
            else:
                d._value = v
        return False


class PendingRun:
    """A queued `CompiledNetwork.run_async`; `result()` waits and returns S[nS,P,P,nF] (same checks as `run`)."""

    def __init__(self, cn, handle, S, done, call):
        self._cn, self._handle, self._S, self._done, self._call = cn, handle, S, done, call

    def result(self) -> np.ndarray:
        if self._done is None:
            status = self._handle.wait()
            if self._cn._hub and (status & HVB_STATUS_ZERO_PIVOT):
                f, drivers, values, f32_out = self._call
                self._done = self._cn.dense_fallback().run(f, drivers, values, f32_out=f32_out)
            else:
                self._cn._raise_for_status(status)
                self._done = self._S
        return self._done


class CompiledNetwork:
    """A frozen `Network`: stamp table + device plan(s) + the library handle(s).

    Narrow-banded systems that still fit the register kernels (9..32 unknowns) get two variants: the
    dense register-kernel plan (`plan` / `handle`, the default) and a banded one (`band_plan` /
    `band_handle`, built on first use).  The banded kernel runs one thread per point, so it wins on
    large batches -- 2x at 10 unknowns, 7-9x at 24-32 (profiles/README.md) -- and loses on short sweeps
    to the latency of a single point; `run` picks by the number of points of the call."""

    BAND_MIN_POINTS = 49152          # about half a grid of the banded kernel (148 SMs x 5 CTAs x 128 threads)
    MULTI_MIN_POINTS = 131072        # points per extra device from which a call is spread over several GPUs

    @classmethod
    def band_min_points(cls, n: int) -> int:
        """Batch size from which the banded variant is used for an n-unknown system: (latency of one point
        in the banded kernel) x (throughput of the register kernel), measured -- 42 k points at n = 10,
        34 k at 16, 13 k at 24, 11 k at 32 (profiles/README.md)."""
        return cls.BAND_MIN_POINTS if n <= 12 else cls.BAND_MIN_POINTS * 2 // 3 if n <= 20 else cls.BAND_MIN_POINTS // 3

    def __init__(self, net, mode: str | None = None, device: int | None = None):
        """`net`: a heavi_b200 `Network` / `Model`, or a `plan.StampTable` frozen by any front-end (the
        reference's own `Network` through the binding of INTEGRATION.md)."""
        self.table = net if isinstance(net, pl.StampTable) else net.stamp_table()
        self._hub, self._dense_fallback, self._mode_arg, self._device_arg = False, None, mode, device
        mode = mode or os.environ.get("HVB_MODE", "auto")
        self._dump = None
        # chain-like netlists (ladders, filters, cascaded lines: tridiagonal after the bandwidth-reducing
        # order): the library generates a kernel for exactly this netlist (kernel class 3), which serves
        # every batch size
        if os.environ.get("HVB_NO_GEN") != "1" and os.environ.get("HVB_FORCE_BLOCK") != "1" and os.environ.get("HVB_NO_BAND") != "1":
            try:
                cplan = pl.compile_plan(self.table, mode, order="rcm")
            except ValueError:
                cplan = None
            if cplan is not None and cplan.band == 1 and cplan.epi_simple and not len(cplan.coops):
                handle = DevicePlanHandle(cplan, device, kernel_hint=2)
                if handle.kernel_info()["kernel_class"] == 3:
                    self.plan, self.handle, self.rows_per_lane, self._band = cplan, handle, 1, False
                    return
            # hubs -- one N-port S block (a Touchstone device) with a matching chain on each of its nodes: the
            # generated kernel E solves two Pn x Pn systems per point instead of one dense (chains + Pn)^2 one
            hplan = pl.compile_plan(self.table, mode)
            if len(hplan.coops) == 1 and hplan.epi_simple and hplan.P <= 8 and int(hplan.coops[0][2]) <= 8 and hplan.n == hplan.n_real:
                handle = DevicePlanHandle(hplan, device, kernel_hint=2)
                ki = handle.kernel_info()
                if ki["kernel_class"] == 3 and ki["gen_hub"] == 1:
                    self.plan, self.handle, self.rows_per_lane, self._band = hplan, handle, 1, False
                    self._hub = True
                    return
        plan = pl.compile_plan(self.table, mode)
        shape = warp_shape(plan.n_real, plan.P, plan.max_nport)
        if shape is not None and shape != (plan.n, plan.ncols - plan.n):
            plan = pl.compile_plan(self.table, plan.mode, pad_to=shape)
        self.plan = plan
        self.handle = DevicePlanHandle(self.plan, device, kernel_hint=1 if shape is not None else 0)
        self.rows_per_lane = self.handle.kernel_info()["rows_per_lane"]
        self._band = None            # (plan, handle) | False once known not to exist
        if shape is None or os.environ.get("HVB_NO_BAND") == "1" or plan.n_real < pl.BAND_SMALL_MIN_ROWS:
            self._band = False

    def _band_variant(self):
        if self._band is None:
            bplan = pl.compile_plan(self.table, self.plan.mode, order="rcm")
            if bplan.band < 0 or bplan.band > 4:
                self._band = False
            else:
                handle = DevicePlanHandle(bplan, self.handle.device, kernel_hint=2)
                self._band = (bplan, handle) if handle.kernel_info()["kernel_class"] == 2 else False
        return self._band

    def variant(self, points: int):
        """(plan, handle) that serves a call of `points` (sample, frequency) points."""
        if self._band is not False and points >= self.band_min_points(self.plan.n_real):
            band = self._band_variant()
            if band:
                return band
        return self.plan, self.handle

    def dense_fallback(self):
        """The same netlist on the precompiled dense kernels (built on first use)."""
        if self._dense_fallback is None:
            old = os.environ.get("HVB_NO_GEN")
            os.environ["HVB_NO_GEN"] = "1"
            try:
                self._dense_fallback = CompiledNetwork(self.table, self._mode_arg, self._device_arg)
            finally:
                if old is None:
                    os.environ.pop("HVB_NO_GEN", None)
                else:
                    os.environ["HVB_NO_GEN"] = old
        return self._dense_fallback

    def replicas(self, handle, points: int):
        """Handles of `handle`'s plan on the other visible devices this call should use (created on first use)."""
        devs = visible_devices()
        if len(devs) < 2 or handle.device != devs[0]:
            return []
        want = min(len(devs), max(1, points // self.MULTI_MIN_POINTS))
        if want < 2:
            return []
        cache = self.__dict__.setdefault("_replica_cache", {})
        out = []
        for d in devs[1:want]:
            key = (id(handle), d)
            if key not in cache:
                cache[key] = DevicePlanHandle(handle.plan, d, kernel_hint=handle.kernel_hint)
            out.append(cache[key])
        return out

    def resolve(self, f, drivers=None, values=None, plan=None):
        return pl.resolve_run(plan or self.plan, f, drivers, values)

    def _tables_follow_drivers(self, f: np.ndarray, drivers, values) -> bool:
        """True when a host-tabulated callable reads one of the batch drivers (e.g. a closure over
        `Random` objects, lib/mc.py RandomTwoPort): such a table differs from sample to sample, which the
        one-table-per-run layout cannot express.  Probed by evaluating every tabulated callable on a few
        frequencies with the drivers set to two different rows of `values`."""
        if not self.plan.table_params or values is None or len(values) < 2 or not len(drivers):
            return False
        fp = np.ascontiguousarray(f[:: max(1, len(f) // 3)][:4], dtype=np.float64)
        rows = []
        with _DriversSet(drivers) as ds:
            for row in (values[0], values[-1], values[len(values) // 2]):
                ds.set(row)
                rows.append([np.array(fn(fp), dtype=np.complex128, copy=True) * np.ones(fp.shape) for (_, _, fn) in self.plan.table_params])
        return any(not np.array_equal(a, b) for other in rows[1:] for a, b in zip(rows[0], other))

    def run(self, f: np.ndarray, drivers=None, values=None, out=None, f32_out=None,
            batch_points: int | None = None, out_ptr: int | None = None, out_ld: int = 0) -> np.ndarray:
        """S[nS,P,P,nF] complex128 on pinned host memory.  `batch_points`: size of the whole job when this
        call is one shard of it, so that every shard picks the same kernel variant (and the result does
        not depend on how the job was cut)."""
        if drivers is not None and self._tables_follow_drivers(f, drivers, values):
            # sample-dependent callables: one run per sample with the drivers set, like the reference's loop
            if out_ptr is not None:
                raise NotImplementedError("sample-dependent callables cannot write into a caller-owned slab (out_ptr)")
            values = np.asarray(values, dtype=np.float64)
            nS = values.shape[0]
            S = out if out is not None else pinned_empty((nS, self.plan.P, self.plan.P, len(f)))
            with _DriversSet(drivers) as ds:
                for s in range(nS):
                    ds.set(values[s])
                    self.run(f, list(drivers), values[s:s + 1], out=S[s:s + 1])
            if f32_out is not None:
                f32_out[:] = f
            return S
        points = len(f) * (1 if values is None else len(values))
        plan, handle = self.variant(points if batch_points is None else batch_points)
        if drivers is not None and values is not None and len(values) == 1 and plan.table_params:
            # one row: host-tabulated callables must see this row's driver values while they are evaluated
            with _DriversSet(drivers) as ds:
                ds.set(np.asarray(values, dtype=np.float64)[0])
                consts, prefs, tables, samples = self.resolve(f, drivers, values, plan)
        else:
            consts, prefs, tables, samples = self.resolve(f, drivers, values, plan)
        reps = self.replicas(handle, points) if (out_ptr is None and batch_points is None) else []
        try:
            S, status = handle.run_host(f, consts, prefs, tables, samples, out, f32_out, out_ptr=out_ptr, out_ld=out_ld, replicas=reps)
        except (RuntimeError, NotImplementedError) as exc:
            # the generated kernels need NVRTC at the first run; if it cannot be loaded or refuses the source, the
            # netlist is served by the library's precompiled kernels instead (still on the GPU) -- said once, loudly
            if handle.kernel_info()["kernel_class"] != 3 or not any(k in str(exc) for k in ("NVRTC", "nvrtc", "kernel generator", "driver entry point", "cuModule")):
                raise
            warnings.warn("heavi_b200: run-time kernel generation unavailable (%s); using the precompiled kernels" % exc, RuntimeWarning)
            fb = self.dense_fallback()
            self.plan, self.handle, self.rows_per_lane, self._band, self._hub = fb.plan, fb.handle, fb.rows_per_lane, fb._band, False
            return self.run(f, drivers, values, out=out, f32_out=f32_out, batch_points=batch_points, out_ptr=out_ptr, out_ld=out_ld)
        if self._hub and (status & HVB_STATUS_ZERO_PIVOT):
            # kernel E does not exchange a chain's last row with its core row; it flags the (rare) point where that
            # pivot is unsafe and the whole call is repeated on the dense register kernel (full partial pivoting)
            return self.dense_fallback().run(f, drivers, values, out=out, f32_out=f32_out, batch_points=batch_points,
                                             out_ptr=out_ptr, out_ld=out_ld)
        if status & HVB_STATUS_NONFINITE:
            # the reference's numba lstsq raises for NaN/Inf input
            raise np.linalg.LinAlgError("MNA matrix or S-parameters contain NaN/Inf (e.g. f = 0 with an inductor)")
        if status & HVB_STATUS_ZERO_PIVOT:
            warnings.warn("singular MNA system at one or more points: the reference returns the SVD "
                          "minimum-norm solution there, the LU solver cannot", RuntimeWarning)
        return S

    def run_async(self, f: np.ndarray, drivers=None, values=None, f32_out=None):
        """`run` without the wait: returns a `PendingRun`; several compiled networks can have one in flight each
        (their kernels and copies share the GPU and the PCIe link back to back)."""
        if drivers is not None and self._tables_follow_drivers(f, drivers, values):
            return PendingRun(self, None, None, self.run(f, drivers, values, f32_out=f32_out), (f, drivers, values, f32_out))
        points = len(f) * (1 if values is None else len(values))
        plan, handle = self.variant(points)
        consts, prefs, tables, samples = self.resolve(f, drivers, values, plan)
        S = handle.run_host_async(f, consts, prefs, tables, samples, f32_out=f32_out)
        return PendingRun(self, handle, S, None, (f, drivers, values, f32_out))

    def run_resident(self, f, drivers=None, values=None, out=None, check: bool = True):
        """Like `run`, but the result stays in HBM: returns a torch CUDA tensor S[nS,P,P,nF] (complex128).
        `f` may be a numpy array or a float64 CUDA tensor; nothing crosses PCIe except f, the sample matrix
        and, with `check`, the 4-byte status word.  For post-processing on the GPU (cost functions of an
        optimiser, statistics, plotting decimation) the 64 B per solve of a two-port never leave the device."""
        import torch
        dev = torch.device("cuda", self.handle.device)
        if isinstance(f, torch.Tensor):
            d_f = f.to(device=dev, dtype=torch.float64).contiguous()
            f_host = d_f.cpu().numpy() if self.plan.table_params else None
        else:
            f_host = np.ascontiguousarray(f, dtype=np.float64)
            d_f = torch.from_numpy(f_host).to(dev)
        nF = int(d_f.shape[0])
        if drivers is not None and self._tables_follow_drivers(f_host if f_host is not None else d_f.cpu().numpy(), drivers, values):
            raise NotImplementedError("sample-dependent callables need the per-trial host path (CompiledNetwork.run)")
        nS = 1 if values is None else len(values)
        plan, handle = self.variant(nF * nS)
        if plan.table_params and f_host is None:
            f_host = d_f.cpu().numpy()
        consts, prefs, tables, samples = pl.resolve_run(plan, f_host if f_host is not None else np.zeros(nF), drivers, values)
        d_tab = torch.from_numpy(tables).to(dev) if tables.shape[0] else None
        d_smp = torch.from_numpy(np.ascontiguousarray(samples)).to(dev)
        if out is None:
            out = torch.empty((nS, plan.P, plan.P, nF), dtype=torch.complex128, device=dev)
        d_status = torch.zeros(1, dtype=torch.int32, device=dev)
        handle.run_device(d_f, consts, prefs, d_tab, d_smp, out, d_status)
        if check:
            status = int(d_status.item())
            if status & HVB_STATUS_RERUN:
                # the two-sided chain kernel met a point it cannot serve safely: this plan goes back to the one-sided kernel
                handle.set_tunable("gen_two_sided", 0)
                d_status.zero_()
                handle.run_device(d_f, consts, prefs, d_tab, d_smp, out, d_status)
                status = int(d_status.item())
            self._raise_for_status(status)
        return out

    def _raise_for_status(self, status: int):
        if status & HVB_STATUS_NONFINITE:
            raise np.linalg.LinAlgError("MNA matrix or S-parameters contain NaN/Inf (e.g. f = 0 with an inductor)")
        if status & HVB_STATUS_ZERO_PIVOT:
            warnings.warn("singular MNA system at one or more points: the reference returns the SVD "
                          "minimum-norm solution there, the LU solver cannot", RuntimeWarning)

    def run_stats(self, f: np.ndarray, drivers, values) -> np.ndarray:
        """Moments over the sample axis, reduced on the device: float64 [6, P, P, nF] in the order
        amplitude mean/std, power mean/std, rms mean/std (StatSparam, sparam.py:215-243)."""
        if self._tables_follow_drivers(f, drivers, values):
            S = self.run(f, drivers, values)                  # per-sample runs; moments taken on the host
            a = np.abs(S)
            pw = a ** 2
            return np.stack([a.mean(axis=0), a.std(axis=0), pw.mean(axis=0), pw.std(axis=0),
                             np.sqrt(pw).mean(axis=0), np.sqrt(pw).std(axis=0)])
        plan, handle = self.variant(len(f) * len(values))
        consts, prefs, tables, samples = self.resolve(f, drivers, values, plan)
        reps = self.replicas(handle, len(f) * len(values))
        if reps and samples.shape[0] >= 1 + len(reps):
            # sample blocks on several devices, each reduced on its device; the [6,P,P,nF] moments of the
            # blocks are merged on the host (counts, means, centred second moments: stable for any split)
            from concurrent.futures import ThreadPoolExecutor
            from .sharding import merge_moments, split_range
            hs = [handle] + reps
            blocks = [split_range(samples.shape[0], len(hs), r) for r in range(len(hs))]
            with ThreadPoolExecutor(len(hs)) as ex:        # ctypes releases the GIL for the duration of each call
                res = list(ex.map(lambda hb: hb[0].run_stats_host(f, consts, prefs, tables, samples[hb[1][0]:hb[1][1]]), zip(hs, blocks)))
            status = 0
            for _, st in res:
                status |= st
            self._raise_for_status(status)
            return merge_moments([r[0] for r in res], [hi - lo for lo, hi in blocks])
        out, status = handle.run_stats_host(f, consts, prefs, tables, samples)
        self._raise_for_status(status)
        return out

    def assemble(self, f: np.ndarray, drivers=None, values=None) -> np.ndarray:
        """Dense reference-layout MNA tensor [(N+M), (N+M), nF] of the first sample (debug/parity)."""
        if self._dump is None:
            dplan = pl.compile_plan(self.table, "dump")
            self._dump = DevicePlanHandle(dplan, self.handle.device)
        dp = self._dump.plan
        f = np.ascontiguousarray(np.asarray(f, dtype=np.float64))
        consts, prefs, tables, samples = pl.resolve_run(dp, f, drivers, values)
        W = self._dump.assemble_host(f, consts, prefs, tables, samples[:1])
        return np.ascontiguousarray(np.moveaxis(W[0], 0, 2))
