"""CPU driver of tests/test_host_logic.py (run in a subprocess: the reference chain's threads never join).
The host adapters (fun::b200_rx, fun::b200_receiver_chain), compiled unchanged against the CPU test double of the C ABI
(fake_b200rx.cpp -> the reference blocks of oracle/_ref), against the reference's own receiver_chain / hot-path blocks on
the same chunked streams.  Prints one JSON object."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from oracle import bind  # noqa: E402

os.environ["B200RX_FAKE_REF"] = bind.REF_SO
FAKE_HOST = os.path.join(HERE, "_build", "libb200host_fake.so")

from stress_receive import make_stream  # noqa: E402
from test_gpu_block import Block, _stream  # noqa: E402
from test_gpu_chain import Chain, _reference_chain  # noqa: E402


def main():
    ref = bind.ref()
    out = {"chain": [], "block": []}
    rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 11)
    n_streams = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    for s in range(n_streams):
        x, snr, nf = make_stream(ref, rng)
        chunk = int(rng.choice([1000, 4096, 4096, 20000]))
        want = _reference_chain(ref, x, chunk)
        ch = Chain(max_frames=256, lib_path=FAKE_HOST)
        got = []
        for pos in range(0, len(x), chunk):
            got += ch.process(x[pos: pos + chunk])
        for _ in range(8):
            got += ch.process(np.zeros(chunk, complex))
        c = ch.counters()
        ch.close()
        # the same stream from a pinned caller buffer in long calls (copied to the GPU straight from it: no staging copy)
        big = int(rng.choice([65536, 70001, 100000]))
        want_big = _reference_chain(ref, x, big)
        chp = Chain(max_frames=256, lib_path=FAKE_HOST)
        got_big, _, _ = chp.run(x, big, pinned=True, max_out=512)
        chp.close()
        assert got_big == want_big, ("pinned direct path", s, big, len(got_big), len(want_big))
        out["chain"].append({"stream": s, "samples": len(x), "chunk": chunk, "snr": snr, "reference": [len(p) for p in want],
                             "adapter": [len(p) for p in got], "equal": got == want, "counters": c})
    # the block adapter on a synchronised stream (tests/test_gpu_block.py scenario)
    for snr in (None, 25):
        brng = np.random.default_rng(17 if snr is None else 18)
        rates = [10, 8, 0, 5, 10, 3, 9, 6, 10, 2, 10, 10]
        lengths = [1500, 300, 40, 700, 64, 1000, 1499, 255, 0, 120, 1500, 333]
        x, payloads = _stream(ref, brng, rates, lengths, snr)
        samples, tags = ref.sync(x, chunk=4096)
        want = ref.hotpath_stream(samples, tags, chunk=4096)
        blk = Block(lib_path=FAKE_HOST)
        got = []
        for s in range(0, len(tags), 4096):
            got += blk.work(samples[s: s + 4096], tags[s: s + 4096])
        got += blk.work(np.zeros(1, complex), np.zeros(1, np.uint8), flush=True)
        c = blk.counters()
        blk.close()
        out["block"].append({"snr": snr, "reference": len(want), "adapter": len(got), "equal": got == want, "counters": c})
    # BASELINE config 1: the test_sim loopback (examples/test_sim.cpp:43-104) through the chain adapter
    data = b"I'm a little tea pot, short and stout.....here is my handle.....blah blah blah.....this rhyme sucks!"
    payload = data * 15
    frame = ref.build_frame(payload, 8)
    x = np.concatenate([np.tile(frame, 40), np.zeros(10 * len(frame), complex)])
    want = _reference_chain(ref, x, 4096)
    ch = Chain(max_frames=256, lib_path=FAKE_HOST)
    got = []
    for pos in range(0, len(x), 4096):
        got += ch.process(x[pos: pos + 4096])
    got += ch.process(None)
    ch.close()
    out["config1"] = {"reference": len(want), "adapter": len(got), "equal": got == want,
                      "all_equal_transmitted": all(g == payload for g in got)}
    # fun::b200_receiver (callback + thread + pause / resume, receiver.cpp:42-77) on the same loopback stream
    import ctypes as C
    lib = C.CDLL(FAKE_HOST)
    lib.b200host_receiver_run.restype = C.c_int
    lib.b200host_receiver_run.argtypes = [C.c_void_p, C.c_long, C.c_long, C.c_long, C.c_int, C.c_uint, C.c_uint, C.c_void_p, C.c_int,
                                          C.c_void_p, C.c_int, C.c_void_p]
    iq = np.ascontiguousarray(x).view(np.float64)
    pl = np.zeros((256, 1500), np.uint8)
    ln = np.zeros(256, np.int32)
    paused = C.c_long(-5)
    n = lib.b200host_receiver_run(iq.ctypes.data, len(x), 4096, 6, 30, 256, 1500, pl.ctypes.data, 1500, ln.ctypes.data, 256,
                                  C.byref(paused))
    rx_got = [bytes(pl[i, : ln[i]]) for i in range(max(n, 0))]
    out["receiver"] = {"payloads": n, "equal": rx_got == want, "rounds_while_paused": int(paused.value)}
    # the block adapter on random streams tagged by the reference's own frame_detector + timing_sync
    out["block_random"] = []
    for s in range(n_streams):
        x, snr, nf = make_stream(ref, rng)
        chunk = int(rng.choice([1000, 4096, 4096, 20000]))
        samples, tags = ref.sync(x, chunk=chunk)
        # whole chunks only: a call too short to complete a symbol vector leaves the reference's frame_decoder without
        # input, and its work() then returns the previous call's payloads again (frame_decoder.cpp:47-48 returns before
        # output_buffer.resize(0)) - a double delivery the adapters do not reproduce
        n_whole = (len(tags) // chunk) * chunk
        samples, tags = samples[:n_whole], tags[:n_whole]
        want = ref.hotpath_stream(samples, tags, chunk=chunk)
        blk = Block(lib_path=FAKE_HOST)
        got = []
        for p in range(0, len(tags), chunk):
            got += blk.work(samples[p: p + chunk], tags[p: p + chunk])
        got += blk.work(np.zeros(1, complex), np.zeros(1, np.uint8), flush=True)
        blk.close()
        out["block_random"].append({"stream": s, "chunk": chunk, "snr": snr, "reference": [len(p) for p in want],
                                    "adapter": [len(p) for p in got], "equal": got == want})
    print(json.dumps(out), flush=True)
    os._exit(0)


if __name__ == "__main__":
    main()
