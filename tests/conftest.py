import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def load_oracle():
    """tests/ is one of the few places allowed to import oracle/ (test infrastructure)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("smer_oracle", os.path.join(ROOT, "oracle", "smer_oracle.py"))
    mod = sys.modules.get("smer_oracle")
    if mod is None:
        mod = importlib.util.module_from_spec(spec)
        sys.modules["smer_oracle"] = mod
        spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def oracle():
    return load_oracle()
