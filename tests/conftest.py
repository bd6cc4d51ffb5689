import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _gpu_run(config):
    """True when the GPU tests are selected (`-m gpu`): there a missing checker is a failure, never a skip."""
    expr = config.getoption("-m") or ""
    return "gpu" in expr and "not gpu" not in expr


@pytest.fixture(scope="session")
def ref(request):
    """The unmodified reference compiled under oracle/_ref (checker).  The library is built here (where /root/reference
    exists) and travels to the GPU box with the repository snapshot; a GPU run without it cannot prove parity, so it
    fails instead of quietly dropping to the committed fixtures."""
    from oracle import bind
    if not bind.have_ref():
        msg = "oracle/_ref/libfunref.so not built (needs /root/reference; run `make -C oracle ref`)"
        if _gpu_run(request.config):
            pytest.fail("parity checker missing under -m gpu: " + msg, pytrace=False)
        pytest.skip(msg)
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
