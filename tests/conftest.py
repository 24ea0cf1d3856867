import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    import oracle
    oracle.build()
    return oracle.Oracle()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference build (oracle/_ref). Skips where it has not been built."""
    import oracle
    if not os.path.exists(oracle.REF_SO):
        if os.path.isdir("/root/reference"):
            oracle.build()
        else:
            pytest.skip("oracle/_ref/libmimc3ref.so not present")
    return oracle.Reference()


@pytest.fixture(scope="session")
def gpu_ctx():
    from mimc3_b200 import lib
    ctx = lib.Context(0)   # raises without a GPU: GPU tests must never pass on a fallback
    yield ctx
    ctx.close()
