import json, os, sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")

@pytest.fixture(scope="session")
def fixtures():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "ref_fixtures.json")))

@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    return oracle_lib.load()

@pytest.fixture(scope="session")
def gbls():
    """The product: ctypes binding over libhbls.so, initialised on cuda:0 (fails loudly without a GPU)."""
    from harmony_b200 import bls
    bls.Init()
    return bls
