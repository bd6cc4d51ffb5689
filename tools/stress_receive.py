#!/usr/bin/env python
"""Randomised parity stress of the raw-capture path: streams with random gaps (including none), rates, lengths, SNRs,
frames cut short, overlapping frames and loud noise bursts go through (a) the reference's whole receiver_chain on the CPU
(oracle/_ref), (b) b200rx_receive on one capture, (c) fun::b200_receiver_chain in random-sized chunks.  Payload sequences
must be identical.  Prints one line per stream and a summary; exit code 1 on any mismatch.

    python tools/stress_receive.py [n_streams] [seed]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fun_ofdm_b200 as fo  # noqa: E402
from oracle import bind  # noqa: E402
from test_gpu_chain import Chain, _reference_chain  # noqa: E402


def make_stream(ref, rng):
    n_frames = int(rng.integers(3, 14))
    snr = float(rng.uniform(8, 32))
    sig = np.sqrt(0.0127 / 10 ** (snr / 10.0) / 2.0)
    parts = [np.zeros(int(rng.integers(0, 600)), complex)]
    for _ in range(n_frames):
        rate = int(rng.integers(0, 11))
        length = int(rng.choice([0, 1, 14, 100, 400, 1000, 1500, int(rng.integers(0, 1500))]))
        f = ref.build_frame(rng.integers(0, 256, length, dtype=np.uint8).tobytes(), rate)
        kind = rng.random()
        if kind < 0.12:      # cut short
            f = f[: int(rng.integers(200, len(f)))]
        elif kind < 0.2:     # scaled up or down
            f = f * float(rng.choice([0.25, 0.5, 2.0, 4.0]))
        parts.append(f)
        g = rng.random()
        if g < 0.25:
            gap = 0          # back to back
        elif g < 0.5:
            gap = int(rng.integers(1, 200))
        else:
            gap = int(rng.integers(200, 1500))
        parts.append(np.zeros(gap, complex))
        if rng.random() < 0.1:  # loud burst of noise
            parts.append(0.3 * (rng.standard_normal(300) + 1j * rng.standard_normal(300)))
    parts.append(np.zeros(3000, complex))
    x = np.concatenate(parts)
    if rng.random() < 0.15 and len(x) > 9000:  # two transmissions on top of each other
        k = int(rng.integers(2000, len(x) - 6000))
        f = ref.build_frame(rng.integers(0, 256, 300, dtype=np.uint8).tobytes(), int(rng.integers(0, 11)))
        x[k: k + len(f)] += f[: len(x) - k]
    x = x + sig * (rng.standard_normal(len(x)) + 1j * rng.standard_normal(len(x)))
    return x, snr, n_frames


def main():
    n_streams = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 2026
    ref = bind.ref()
    rng = np.random.default_rng(seed)
    rx = fo.Receiver(0, 256, 4095)
    bad = 0
    total = 0
    for s in range(n_streams):
        x, snr, nf = make_stream(ref, rng)
        # (a) one capture against the reference fed the whole stream in one call
        want1 = _reference_chain(ref, x, len(x))
        got1, info = rx.receive(np.concatenate([x, np.zeros(400, complex)]))
        # (b) the chunked GPU chain against the reference chain fed the same chunks
        chunk = int(rng.choice([333, 1000, 4096, 4096, 20000]))
        want2 = _reference_chain(ref, x, chunk)
        ch = Chain(max_frames=256)
        got2 = []
        for pos in range(0, len(x), chunk):
            got2 += ch.process(x[pos: pos + chunk])
        for _ in range(8):  # the same drain the reference gets
            got2 += ch.process(np.zeros(chunk, complex))
        ch.close()
        ok1, ok2 = got1 == want1, got2 == want2
        total += len(want2)
        if not (ok1 and ok2):
            bad += 1
        print("stream %2d: %6d samples, %2d frames sent at %4.1f dB; one call: reference %2d payloads, GPU %2d %s; chunks of %5d: "
              "reference %2d, GPU chain %2d %s" % (s, len(x), nf, snr, len(want1), len(got1), "ok" if ok1 else "MISMATCH", chunk,
                                                  len(want2), len(got2), "ok" if ok2 else "MISMATCH"), flush=True)
        if not ok1:
            print("    one call : reference", [len(p) for p in want1], "GPU", [len(p) for p in got1])
        if not ok2:
            print("    chunked  : reference", [len(p) for p in want2], "GPU", [len(p) for p in got2])
    print("streams %d, reference payloads %d, streams with a mismatch %d" % (n_streams, total, bad))
    sys.stdout.flush()
    os._exit(1 if bad else 0)  # the reference chain's threads never join


if __name__ == "__main__":
    main()
