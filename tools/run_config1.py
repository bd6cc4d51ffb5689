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
    bits = 8 * len(payload)
    points = []
    warm = Chain(max_frames=2048, max_payload=1500)  # CUDA context, module load
    warm.run(np.tile(frame, 4), 4096)
    warm.close()
    # (frames, zero pad in frame lengths, samples per call); the first row is examples/test_sim.cpp as it is: 100 frames
    # followed by 1000 frame lengths of zeros (test_sim.cpp:58-71), 4096-sample calls
    for n_frames, pad, chunk in ((100, 1000, 4096), (100, 10, 4096), (2000, 10, 4096), (8000, 10, 65536), (8000, 10, 1 << 20)):
        x = np.concatenate([np.tile(frame, n_frames), np.zeros(pad * len(frame), complex)])
        chain = ref.chain_new()
        want, t_ref = ref.chain_run(chain, x, chunk, drain=8, max_len=1500)  # native loop; the drain is not timed
        point = {"frames": n_frames, "zero_pad_frame_lengths": pad, "chunk_samples": chunk, "samples": int(len(x)),
                 "reference_chain": {"payloads": len(want), "seconds": t_ref, "mbit_s": len(want) * bits / t_ref / 1e6,
                                     "msamples_s": len(x) / t_ref / 1e6}, "b200_receiver_chain": {}}
        # variants of the same feed loop (native, host_capi.cpp b200host_chain_run; seconds include the final flush):
        #   test_sim      std::vector built per chunk and passed by value, exactly as examples/test_sim.cpp:84-87
        #   pointer       process_samples(const complex<double>*, n) on pageable memory
        #   pinned        the caller's samples already lie in pinned memory (alloc_samples), e.g. a radio driver's buffer
        #   lag           depth / max_lag: passes in flight / calls a payload may stay behind (5 = the reference's own)
        variants = [("test_sim_by_value_depth6_lag5", dict(by_value=True), dict(depth=6, max_lag=5)),
                    ("pointer_depth6_lag5", dict(), dict(depth=6, max_lag=5)),
                    ("pointer_depth12_lag11", dict(), dict(depth=12, max_lag=11)),
                    ("pointer_depth1_synchronous", dict(), dict(depth=1, max_lag=0)),
                    ("pinned_depth6_lag5", dict(pinned=True), dict(depth=6, max_lag=5))]
        for name, run_kw, new_kw in variants:
            g = Chain(max_frames=2048, max_payload=1500, **new_kw)
            # untimed: every lane allocates its staging on first use (cudaMalloc / cudaHostAlloc take milliseconds)
            g.run(x[: min(len(x), 14 * chunk)], chunk, max_out=8192, stride=1500, **run_kw)
            got, t_feed, t_all = g.run(x, chunk, max_out=8192, stride=1500, **run_kw)
            g.close()
            point["b200_receiver_chain"][name] = {
                "payloads": len(got), "seconds": t_all, "seconds_feed_loop": t_feed, "mbit_s": len(got) * bits / t_all / 1e6,
                "msamples_s": len(x) / t_all / 1e6, "us_per_call": 1e6 * t_feed / max(1, -(-len(x) // chunk)),
                "payload_sequence_identical_to_reference": got == want,
                "all_payloads_equal_transmitted": all(p == payload for p in got)}
        points.append(point)
    print(json.dumps({"config": "1: test_sim loopback, RATE_3_4_QAM16, 1500-byte ASCII payload, frames back to back, noiseless; "
                                "reference = its six-thread receiver_chain on %d host cores (FFTW/Boost replaced by "
                                "oracle/shims), fed from Python per chunk; GPU = fun::b200_receiver_chain on one B200, host "
                                "samples in, payloads out, native feed loop" % (os.cpu_count() or 1), "points": points}), flush=True)
    os._exit(0)  # the reference chain's threads never join


if __name__ == "__main__":
    main()
