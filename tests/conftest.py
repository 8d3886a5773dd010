import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_package():
    """The package directory name is not an identifier; load it as module ``dsc_b200``."""
    if "dsc_b200" in sys.modules:
        return sys.modules["dsc_b200"]
    d = os.path.join(ROOT, "3d-image-segmentation-with-mesh_b200")
    spec = importlib.util.spec_from_file_location("dsc_b200", os.path.join(d, "__init__.py"), submodule_search_locations=[d])
    m = importlib.util.module_from_spec(spec)
    sys.modules["dsc_b200"] = m
    spec.loader.exec_module(m)
    return m


@pytest.fixture(scope="session")
def pkg():
    return load_package()


@pytest.fixture(scope="session")
def flat():
    from oracle import flat as f
    f.lib()
    return f
