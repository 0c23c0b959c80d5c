"""Import the UNMODIFIED reference (`/root/reference/src/heavi`) in this container.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this file.  It exists so
that `oracle/make_golden.py` and `oracle/check_oracle_vs_ref.py` can run the real reference
and pin the numpy oracle (`oracle/mna_oracle.py`) against it.  `/root/reference` does not
exist on the GPU box, so nothing under `tests/ -m gpu`, `smoke()` or `bench.py` calls this.

No reference source is copied into the repo.  Three things keep the reference from importing
on this image and all three are handled in memory:

* `numba_progress` is not installed  -> a stub module whose `ProgressBarType` is a real numba
  type (the reference uses it in an eagerly compiled signature, solvers/tranfn.py:76-86).
* `emsutil` (plotting) is not installed -> a stub with no-op `plot`, `smith`, `plot_sp`.
* `heavi/node.py:27-46`: `UnassignedType` defines `__eq__` without `__hash__`, which Python
  >= 3.11 rejects as a dataclass default.  The module source is read from the read-only tree,
  one line (`__hash__ = object.__hash__`) is added to the in-memory text, and the result is
  executed as `heavi.node` before the package `__init__` runs.
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import sys
import types

REF_SRC = "/root/reference/src"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_SRC, "heavi"))


def _install_stubs() -> None:
    if "numba_progress" not in sys.modules:
        from numba import int64
        from numba.experimental import jitclass

        @jitclass([("count", int64)])
        class _PB:
            def __init__(self):
                self.count = 0

            def update(self, n):
                self.count += n

        class ProgressBar:
            def __init__(self, *a, **k):
                self._pb = _PB()

            def __enter__(self):
                return self._pb

            def __exit__(self, *exc):
                return False

        pkg = types.ModuleType("numba_progress")
        sub = types.ModuleType("numba_progress.progress")
        sub.ProgressBar = ProgressBar
        sub.ProgressBarType = _PB.class_type.instance_type
        pkg.ProgressBar = ProgressBar
        pkg.ProgressBarType = sub.ProgressBarType
        pkg.progress = sub
        sys.modules["numba_progress"] = pkg
        sys.modules["numba_progress.progress"] = sub
    if "emsutil" not in sys.modules:
        ems = types.ModuleType("emsutil")
        ems.plot = lambda *a, **k: None
        ems.smith = lambda *a, **k: None
        ems.plot_sp = lambda *a, **k: None
        sys.modules["emsutil"] = ems


def load_reference():
    """Return the reference `heavi` package, imported from the read-only tree."""
    if "heavi" in sys.modules:
        return sys.modules["heavi"]
    if not reference_available():
        raise RuntimeError("reference tree /root/reference is not present on this machine")
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_heavi_ref")
    sys.dont_write_bytecode = True
    _install_stubs()

    pkg_dir = os.path.join(REF_SRC, "heavi")
    spec = importlib.util.spec_from_file_location(
        "heavi", os.path.join(pkg_dir, "__init__.py"), submodule_search_locations=[pkg_dir]
    )
    pkg = importlib.util.module_from_spec(spec)
    sys.modules["heavi"] = pkg

    # heavi.node with the one-line Python>=3.11 dataclass fix, applied to the in-memory text.
    node_path = os.path.join(pkg_dir, "node.py")
    with open(node_path, "r") as fh:
        src = fh.read()
    marker = "UNASSIGNED = UnassignedType()"
    assert marker in src
    src = src.replace(marker, "UnassignedType.__hash__ = object.__hash__\n" + marker, 1)
    node_mod = types.ModuleType("heavi.node")
    node_mod.__file__ = node_path
    node_mod.__package__ = "heavi"
    sys.modules["heavi.node"] = node_mod
    exec(compile(src, node_path, "exec"), node_mod.__dict__)
    pkg.node = node_mod

    spec.loader.exec_module(pkg)
    return pkg


if __name__ == "__main__":
    import time

    t0 = time.time()
    hv = load_reference()
    print("reference heavi imported in %.1f s" % (time.time() - t0), hv.__file__)
