import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # a kernel that never returns must fail the test, not hang the box: with pytest-timeout present every test gets a
    # generous limit unless the command line sets one (the slowest test, the unmodified reference driver making 983 040
    # one-point launches, takes about a minute)
    if config.pluginmanager.hasplugin("timeout") and not getattr(config.option, "timeout", None):
        config.option.timeout = 1500


@pytest.fixture(scope="session")
def oracle():
    from oracle_lib import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reflib():
    from oracle_lib import RefLib
    if not RefLib.available():
        pytest.skip("oracle/_ref/libwnref.so not built (needs /root/reference at build time)")
    return RefLib()


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def ref_vectors(golden_dir):
    import numpy as np
    return dict(np.load(os.path.join(golden_dir, "ref_vectors.npz")))


@pytest.fixture(scope="session")
def tiles128(oracle):
    """The two tiles every reference driver uses: n=128, seed 12345 (experient/main.cpp:136-144)."""
    return {2: oracle.generate_tile(128, 12345, 2), 3: oracle.generate_tile(128, 12345, 3)}
