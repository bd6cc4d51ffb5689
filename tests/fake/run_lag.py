"""CPU driver of tests/test_host_logic.py::test_delivery_lag_rules: the chain adapter against a CPU double whose passes
stay "in flight" (B200RX_FAKE_SLOW_POLLS): a payload must come back no later than max_lag calls after the call that
completed its frame, never earlier than the pass has finished, in stream order, and flush() must return the rest."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import bind  # noqa: E402

os.environ["B200RX_FAKE_REF"] = bind.REF_SO
os.environ["B200RX_FAKE_SLOW_POLLS"] = sys.argv[1] if len(sys.argv) > 1 else "1000000"
FAKE_HOST = os.path.join(HERE, "_build", "libb200host_fake.so")
from test_gpu_chain import Chain  # noqa: E402


def main():
    ref = bind.ref()
    rng = np.random.default_rng(3)
    frames = [ref.build_frame(rng.integers(0, 256, 200, dtype=np.uint8).tobytes(), 10) for _ in range(12)]
    gap = 3000
    x = np.concatenate([np.concatenate([np.zeros(gap, complex), f]) for f in frames] + [np.zeros(40 * 4096, complex)])
    x = x + 1e-4 * (rng.standard_normal(len(x)) + 1j * rng.standard_normal(len(x)))
    ends = np.cumsum([gap + len(f) for f in frames]) - 1   # stream index of each frame's last sample
    chunk = 4096
    out = {}
    for depth, lag in ((6, 5), (4, 2), (8, 7), (1, 0)):
        ch = Chain(max_frames=64, lib_path=FAKE_HOST, depth=depth, max_lag=lag)
        arrival = []
        for c, pos in enumerate(range(0, len(x), chunk)):
            got = ch.process(x[pos: pos + chunk])
            arrival += [c] * len(got)
        tail = ch.process(None)
        ch.close()
        # a frame is complete in the call that brings its last sample - or one later when its STS_END tag waits in the
        # last 160 samples of a call (timing_sync.cpp:68)
        done_call = [int(e // chunk) for e in ends]
        lags = [a - d for a, d in zip(arrival, done_call)]
        out["%d/%d" % (depth, lag)] = {"payloads": len(arrival) + len(tail), "lags": lags, "left_for_flush": len(tail)}
    # the block adapter on the same stream with genie tags (LTS1 at frame start + 184), rounds of 4096 samples
    from test_gpu_block import Block
    tags = np.zeros(len(x), np.uint8)
    starts = np.cumsum([gap + len(f) for f in frames]) - np.array([len(f) for f in frames])
    tags[starts + 184] = 4
    for depth, lag in ((4, 3), (6, 1), (1, 0)):
        blk = Block(max_frames=64, lib_path=FAKE_HOST, depth=depth, max_lag=lag)
        arrival = []
        for c, pos in enumerate(range(0, len(x), chunk)):
            got = blk.work(x[pos: pos + chunk], tags[pos: pos + chunk])
            arrival += [c] * len(got)
        tail = blk.work(np.zeros(1, complex), np.zeros(1, np.uint8), flush=True)
        blk.close()
        # complete when sample LTS1 + 128 + 80 * (1 + nsym) - 1 has arrived: the frame's last sample
        done_call = [int(e // chunk) for e in ends]
        out["block %d/%d" % (depth, lag)] = {"payloads": len(arrival) + len(tail), "lags": [a - d for a, d in zip(arrival, done_call)],
                                            "left_for_flush": len(tail)}
    print(json.dumps(out), flush=True)
    os._exit(0)


if __name__ == "__main__":
    main()
