import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def oracle():
    import bindings
    if not os.path.exists(bindings.ORACLE_SO):
        import __graft_entry__ as g
        g.build_oracle()
    return bindings.Oracle()


@pytest.fixture(scope="session")
def reference():
    import bindings
    if not bindings.Reference.available():
        pytest.skip("oracle/_ref/libtmc2ref.so not built (needs /root/reference)")
    return bindings.Reference()


@pytest.fixture(scope="session")
def product():
    import bindings
    p = bindings.Product(0)
    yield p
    p.close()
