"""pytest configuration: markers and shared loaders.

`-m "not gpu"`: oracle vs golden vectors / compiled reference, host logic (parameters, frame,
plans) through the C-ABI, and that libtspws_b200.so loads and exports every declared symbol.
`-m gpu`: parity of the CUDA path against the oracle, through the C-ABI.
"""
import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_pkg():
    """Import ts-pws_b200/ (the directory name is not a Python identifier)."""
    if "ts_pws_b200" in sys.modules:
        return sys.modules["ts_pws_b200"]
    path = os.path.join(ROOT, "ts-pws_b200", "__init__.py")
    spec = importlib.util.spec_from_file_location("ts_pws_b200", path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ts_pws_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()


@pytest.fixture(scope="session")
def lib(pkg):
    so = pkg.LIB_PATH
    if not os.path.exists(so):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.dirname(so)], stdout=subprocess.DEVNULL)
    return pkg.load_library()


@pytest.fixture(scope="session")
def gpu_lib(lib):
    n = lib.tspws_b200_device_count()
    if n <= 0:
        pytest.fail(f"no CUDA device visible to libtspws_b200.so (device_count={n}); "
                    "gpu tests must run on the GPU box -- there is no CPU fallback")
    return lib
