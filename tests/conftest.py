import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def oracle():
    import sxtest
    return sxtest.load_oracle()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference driver (oracle/_ref/libsx_ref.so); skip when it was never built."""
    import sxtest
    lib = sxtest.load_reference()
    if lib is None:
        pytest.skip("oracle/_ref/libsx_ref.so not built (needs /root/reference)")
    return lib


@pytest.fixture(scope="session")
def ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test run without a GPU")
    from sxxcvr_b200 import Context
    c = Context(0)
    yield c
    c.close()
