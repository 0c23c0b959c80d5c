import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")
    config.addinivalue_line("markers", "library_kernels: GPU test of the precompiled kernels only (A/A2/B/C); "
                                       "the generated chain kernel (class 3) is switched off")


def pytest_generate_tests(metafunc):
    """Every GPU test runs twice: with the generated chain kernels (kernel class 3, the default product
    path for ladders / filters / cascades) and with the library's precompiled kernels only (HVB_NO_GEN=1)."""
    if metafunc.definition.get_closest_marker("gpu") is None or "kernel_path" not in metafunc.fixturenames:
        return
    legacy = metafunc.definition.get_closest_marker("library_kernels") is not None
    metafunc.parametrize("kernel_path", ["library"] if legacy else ["generated", "library"], indirect=True)


@pytest.fixture(autouse=True)
def kernel_path(request):
    mode = getattr(request, "param", None)
    if mode is None:
        yield None
        return
    old = os.environ.get("HVB_NO_GEN")
    if mode == "library":
        os.environ["HVB_NO_GEN"] = "1"
    else:
        os.environ.pop("HVB_NO_GEN", None)
    yield mode
    if old is None:
        os.environ.pop("HVB_NO_GEN", None)
    else:
        os.environ["HVB_NO_GEN"] = old


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
