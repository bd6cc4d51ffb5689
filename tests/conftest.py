import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference compiled under oracle/_ref (checker)."""
    from oracle import bind
    if not bind.have_ref():
        pytest.skip("oracle/_ref/libfunref.so not built (needs /root/reference; run `make -C oracle ref`)")
    return bind.ref()


@pytest.fixture(scope="session")
def port():
    """The plain-C restatement oracle/ofdm_oracle.c (checker)."""
    import subprocess
    from oracle import bind
    if not os.path.exists(bind.PORT_SO):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"])
    return bind.port()


@pytest.fixture(scope="session")
def rx_factory():
    """Factory for product handles on cuda:0; fails loudly if the CUDA library is missing."""
    import fun_ofdm_b200 as fo
    made = []

    def make(max_frames=512, max_payload_bytes=1500):
        r = fo.Receiver(0, max_frames, max_payload_bytes)
        made.append(r)
        return r

    yield make
    for r in made:
        r.close()
