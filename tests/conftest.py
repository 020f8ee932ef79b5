import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def load_builder():
    """soundml_b200/build.py loaded by path: importing the package would load the
    shared library, which may be stale or absent before the build."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "soundml_b200_build", os.path.join(ROOT, "soundml_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def lib():
    """The product library, built on demand (nvcc cross-compiles without a GPU)."""
    load_builder().build()
    import soundml_b200
    return soundml_b200


@pytest.fixture(scope="session")
def goldens():
    from golden_util import Goldens
    return Goldens()
