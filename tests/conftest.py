import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_generate_tests(metafunc):
    """GPU parity tests run twice: against the restatement (oracle/*.cpp) and against the REFERENCE'S OWN code
    (oracle/_ref/libviorb_ref.so, built by oracle/refbuild from /root/reference unmodified; the prebuilt library travels
    to the GPU box).  CPU tests use the restatement; tests/test_ref_*.py compare the two directly."""
    if "checker_kind" in metafunc.fixturenames:
        gpu = metafunc.definition.get_closest_marker("gpu") is not None
        metafunc.parametrize("checker_kind", ["restatement", "reference"] if gpu else ["restatement"], scope="session")


@pytest.fixture(scope="session")
def oracle(checker_kind):
    from oracle import oracle_py
    oracle_py.lib()
    if checker_kind == "restatement":
        yield oracle_py
        return
    from oracle import ref_py
    if not ref_py.available():
        pytest.skip("no reference library (oracle/_ref/libviorb_ref.so) and no /root/reference to build it from")
    ref_py.set_allocator(1)
    with oracle_py.using(ref_py.lib(fallback=True)):
        yield oracle_py


@pytest.fixture(scope="session")
def pattern():
    from util import load_pattern
    return load_pattern()
