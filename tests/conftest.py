"""Test configuration.  `-m "not gpu"` covers the oracle, the host logic and the C-ABI export surface
(no GPU needed); `-m gpu` holds the parity tests proper, which call the CUDA path through the C ABI."""
import os
import sys

import pytest

# the one-GPU tile-split tests run up to 8 engines ("ranks") on one device, each spinning briefly on its peers: give
# every stream its own hardware queue so that no rank queues behind a waiting one (read at CUDA initialisation)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gs-evt_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run on the GPU box only")


@pytest.fixture(scope="session")
def built():
    """Makes sure libgsevt.so and the CPU oracle exist (compiles them when missing)."""
    import __graft_entry__ as ge
    from gsevt import lib
    if not os.path.exists(lib.LIB_PATH) or not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        ge.build()
    return lib.load()


@pytest.fixture(scope="session")
def cuda_dev(built):
    import torch
    from gsevt import lib
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test selected but no CUDA device is visible (there is no CPU fallback)")
    lib.require_device()
    return torch.device("cuda:0")


def load_reference_extension():
    """The unmodified reference operator built by oracle/build_ref.sh (None when absent)."""
    p = os.path.join(ROOT, "oracle", "_ref", "ext", "diff_gaussian_rasterization")
    if not os.path.isdir(p) or not any(f.startswith("_C") and f.endswith(".so") for f in os.listdir(p)):
        return None
    if "ref_dgr" in sys.modules:
        return sys.modules["ref_dgr"]
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_dgr", os.path.join(p, "__init__.py"), submodule_search_locations=[p])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_dgr"] = mod
    spec.loader.exec_module(mod)
    return mod
