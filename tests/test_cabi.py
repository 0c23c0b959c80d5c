"""CPU: the C-ABI library loads and exports every symbol include/heavi_b200.h declares.
No compute call is made (there is no GPU in the build container)."""
import ctypes
import os
import re

from heavi_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "heavi_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hvb_[a-z_0-9]+)\s*\(", text)))


def test_header_functions_are_exported():
    lib = engine.load_library()
    names = declared_functions()
    assert set(names) >= set(engine.EXPORTS)
    for n in names:
        assert hasattr(lib, n), n


def test_version_and_error_strings():
    lib = engine.load_library()
    assert b"sm_100a" in lib.hvb_version()
    assert isinstance(lib.hvb_last_error(), bytes)


def test_struct_layouts_match_header():
    # 18 int32 + 19 pointers; 8-byte aligned
    assert ctypes.sizeof(engine.PlanDesc) == 18 * 4 + 19 * 8
    assert ctypes.sizeof(engine.RunArgs) == 10 * 8
    assert ctypes.sizeof(engine.KernelInfo) == 10 * 4 + 8 + 4 * 4 + 8 + 8
    assert ctypes.sizeof(engine.CodegenInfo) == 4 * 4 + 5 * 8 + 2 * 4 + 8 + 2 * 4 + 256


def test_plan_format_constants_in_sync():
    from heavi_b200 import plan as pl
    text = open(os.path.join(ROOT, "heavi_b200", "csrc", "plan_format.h")).read()
    defs = dict(re.findall(r"#define\s+(\w+)\s+\(?(-?\d+)\)?\s", text))
    for name in ("PK_CONST", "PK_SAMPLE", "PK_SAMPLE_NEG", "PK_SAMPLE_INV", "PK_TABLE", "PK_SPLINE", "PK_BETA",
                 "OP_COPY", "OP_RECIP", "OP_CAP", "OP_IND", "OP_TLINE", "COOP_NPORT_S", "COOP_NPORT_Y",
                 "ROW_GROUND", "ROW_VIRTUAL_INT"):
        assert int(defs[name]) == getattr(pl, name), name
    assert int(defs["HVB_OP_WORDS"]) == pl.OP_WORDS and int(defs["HVB_COOP_WORDS"]) == pl.COOP_WORDS
    assert int(defs["HVB_MAX_NPORT"]) == pl.MAX_NPORT
