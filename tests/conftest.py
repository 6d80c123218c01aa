"""pytest configuration: `gpu` marker, package loader (the package dir has '-' in its name)."""
import importlib.util
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG_DIR = ROOT / "structure-plp-slam_b200"
sys.path.insert(0, str(ROOT / "tests"))


def load_package():
    """Import structure-plp-slam_b200/ as module `plpslam_b200` (by path)."""
    if "plpslam_b200" in sys.modules:
        return sys.modules["plpslam_b200"]
    spec = importlib.util.spec_from_file_location("plpslam_b200", PKG_DIR / "__init__.py",
                                                  submodule_search_locations=[str(PKG_DIR)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["plpslam_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def plp():
    return load_package()


@pytest.fixture(scope="session")
def ctx(plp):
    c = plp.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def orc():
    import oracle_api
    return oracle_api.Oracle()
