import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
CACHE = os.path.join(ROOT, "scenes", "_cache")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: minutes of CPU oracle work (runs with TB_RUN_SLOW=1)")


def scene_path(name):
    """Bundled scene as .tbscene: committed golden copy first, then the build-time cache."""
    for d in (GOLDEN, CACHE):
        p = os.path.join(d, name + ".tbscene")
        if os.path.exists(p):
            return p
    return None


@pytest.fixture(scope="session")
def built():
    """Make sure the native libraries exist (builds them here on the CPU box; prebuilt on the GPU box)."""
    from tracerboy_b200 import build, lib_path
    from oracle import binding
    if not os.path.exists(lib_path()):
        build.build_product()
    if not os.path.exists(binding.lib_path()):
        build.build_oracle()
    return True


@pytest.fixture(scope="session")
def cornell(built):
    p = scene_path("cornell-box")
    assert p, "cornell-box.tbscene missing (tests/golden)"
    return p


@pytest.fixture(scope="session")
def teapot(built):
    p = scene_path("teapot")
    if not p:
        pytest.skip("teapot.tbscene not in scenes/_cache (needs the reference mount at build time)")
    return p


@pytest.fixture(scope="session")
def vwvan(built):
    p = scene_path("vw-van")
    if not p:
        pytest.skip("vw-van.tbscene not in scenes/_cache (needs the reference mount at build time)")
    return p


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
