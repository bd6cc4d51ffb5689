"""CPU tests of the host logic: the two adapters that sit between fun_ofdm's API and the C ABI - fun::b200_rx (tag state
machine, frame cutting, batching) and fun::b200_receiver_chain (streaming state across process_samples calls, chunk
boundaries, deduplication) - compiled UNCHANGED against a CPU test double of libb200rx.so (tests/fake/fake_b200rx.cpp: the
reference's own blocks behind the same entry points) and compared with the reference's receiver_chain / hot-path blocks on
the same chunked streams.  The GPU tests run the same adapters against the real library."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FAKE = os.path.join(ROOT, "tests", "fake")


@pytest.fixture(scope="module")
def fake_built(ref):
    import shutil
    cxx = "/opt/gcc/bin/g++" if os.path.exists("/opt/gcc/bin/g++") else (shutil.which("g++") or "g++")
    subprocess.check_call(["make", "-C", FAKE, "CXX=" + cxx], stdout=subprocess.DEVNULL)
    return os.path.join(FAKE, "_build", "libb200host_fake.so")


def _run(seed, n_streams):
    r = subprocess.run([sys.executable, os.path.join(FAKE, "run_fake.py"), str(seed), str(n_streams)], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_adapters_match_reference_on_chunked_streams(fake_built, seed):
    out = _run(seed, 12)
    assert len(out["chain"]) == 12
    delivered = 0
    for c in out["chain"]:
        assert c["equal"], c
        delivered += len(c["reference"])
    assert delivered >= 30, delivered
    for b in out["block"]:
        assert b["equal"] and b["adapter"] >= 11, b
    assert out["config1"]["equal"] and out["config1"]["adapter"] == 40 and out["config1"]["all_equal_transmitted"], out["config1"]
    # fun::b200_receiver: same payload sequence through the callback; no round runs while the receiver is paused
    assert out["receiver"]["equal"] and out["receiver"]["payloads"] == 40 and out["receiver"]["rounds_while_paused"] == 0, out["receiver"]
    assert len(out["block_random"]) == 12
    for b in out["block_random"]:
        assert b["equal"], b


def test_delivery_lag_rules(fake_built):
    """fun::b200_receiver_chain hands a payload out as soon as its pass has finished and never later than max_lag calls
    after the call that completed the frame (the reference's own chain: five).  A CPU double whose passes never finish
    on their own shows the bound (lag == max_lag for every frame, 0 for the synchronous chain), one whose passes finish
    at once shows the floor (the next call); nothing is lost, flush() has nothing left to return."""
    def run(slow_polls):
        r = subprocess.run([sys.executable, os.path.join(FAKE, "run_lag.py"), str(slow_polls)], capture_output=True, text=True,
                           timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        return json.loads(r.stdout.strip().splitlines()[-1])
    never = run(10 ** 9)
    for key, lag in (("6/5", 5), ("4/2", 2), ("8/7", 7), ("1/0", 0)):
        assert never[key]["payloads"] == 12 and never[key]["left_for_flush"] == 0, never[key]
        assert never[key]["lags"] == [lag] * 12, (key, never[key]["lags"])
    for key, lag in (("block 4/3", 3), ("block 6/1", 1), ("block 1/0", 0)):  # fun::b200_rx: rounds instead of calls
        assert never[key]["payloads"] == 12 and never[key]["left_for_flush"] == 0, never[key]
        assert never[key]["lags"] == [lag] * 12, (key, never[key]["lags"])
    at_once = run(0)
    for key in ("6/5", "4/2", "8/7"):
        assert at_once[key]["payloads"] == 12 and set(at_once[key]["lags"]) <= {1}, (key, at_once[key])
    assert at_once["1/0"]["lags"] == [0] * 12


def test_fake_library_is_test_only():
    """the test double must never be reachable from the product tree"""
    for base, _, files in os.walk(os.path.join(ROOT, "fun_ofdm_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".h", ".cu", ".cuh", "Makefile")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "fake_b200rx" not in text and "b200rx_fake" not in text, f
    assert not os.path.exists(os.path.join(ROOT, "fun_ofdm_b200", "lib", "libb200rx_fake.so"))
