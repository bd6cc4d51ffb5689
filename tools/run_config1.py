#!/usr/bin/env python
"""BASELINE.json configs[0]: the examples/test_sim loopback (frame_builder -> receiver_chain, no USRP; RATE_3_4_QAM16, the
100-character string x 15 = 1500 bytes, identical frames back to back, noiseless, zero pad) - through the reference's own
six-thread receiver_chain on the host CPU (oracle/_ref, unmodified sources) and through fun::b200_receiver_chain on one
B200, fed with the same chunks.  test_sim uses 100 frames and 4096-sample chunks; more frames and larger chunks are added
because a 4096-sample call is shorter than one frame and measures call latency, not throughput.

Prints one JSON object: payload sequences must be identical; Mbit/s = CRC-OK payload bits / wall time of the feed loop.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import bind  # noqa: E402
from test_gpu_chain import Chain  # noqa: E402


def feed(process, x, chunk):
    out = []
    t0 = time.perf_counter()
    for s in range(0, len(x), chunk):
        out += process(x[s: s + chunk])
    return out, time.perf_counter() - t0


def main():
    ref = bind.ref()
    data = b"I'm a little tea pot, short and stout.....here is my handle.....blah blah blah.....this rhyme sucks!"
    payload = data * 15
    frame = ref.build_frame(payload, 8)
    points = []
    for n_frames, chunk in ((100, 4096), (2000, 4096), (2000, 65536), (2000, 1 << 20)):
        x = np.concatenate([np.tile(frame, n_frames), np.zeros(10 * len(frame), complex)])
        chain = ref.chain_new()
        want, t_ref = feed(lambda c: ref.chain_process(chain, c, max_frames=2048), x, chunk)
        for _ in range(8):  # drain the reference pipeline (not timed; one round per block)
            want += ref.chain_process(chain, np.zeros(4096, complex), max_frames=2048)
        g = Chain(max_frames=2048, max_payload=1500)
        g.process(x[: 4 * len(frame)], max_out=2048, stride=1500)  # warm-up: CUDA context, staging buffers
        g.close()
        g = Chain(max_frames=2048, max_payload=1500)
        g.process(np.zeros(64, complex))
        got, t_gpu = feed(lambda c: g.process(c, max_out=2048, stride=1500), x, chunk)
        got += g.process(None)
        g.close()
        bits = 8 * len(payload)
        points.append({"frames": n_frames, "chunk_samples": chunk, "samples": int(len(x)),
                       "reference_chain": {"payloads": len(want), "seconds": t_ref, "mbit_s": len(want) * bits / t_ref / 1e6,
                                           "msamples_s": len(x) / t_ref / 1e6},
                       "b200_receiver_chain": {"payloads": len(got), "seconds": t_gpu, "mbit_s": len(got) * bits / t_gpu / 1e6,
                                               "msamples_s": len(x) / t_gpu / 1e6},
                       "payload_sequences_identical": got == want, "all_payloads_equal_transmitted": all(p == payload for p in got)})
    print(json.dumps({"config": "1: test_sim loopback, RATE_3_4_QAM16, 1500-byte ASCII payload, frames back to back, noiseless; "
                                "reference = its six-thread receiver_chain on %d host cores (FFTW/Boost replaced by "
                                "oracle/shims), GPU = fun::b200_receiver_chain on one B200, host std::vector in, payloads out"
                                % (os.cpu_count() or 1), "points": points}), flush=True)
    os._exit(0)  # the reference chain's threads never join


if __name__ == "__main__":
    main()
