"""Shared helpers for the tests: corpus construction and comparison against the checker."""
import numpy as np

LTS1_OFFSET = 184  # timing_sync tags LTS1 8 samples early: frame_start + 160 + 24 (timing_sync.cpp:105)


def awgn(rng, n, sigma):
    return sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))


def snr_sigma(frame, snr_db):
    """Per-component noise sigma for a given SNR over the post-preamble part of the frame."""
    p = np.mean(np.abs(frame[320:]) ** 2)
    return np.sqrt(p / 10 ** (snr_db / 10.0) / 2.0)


def make_corpus(builder, rng, rates, lengths, snr_db=None, gap=64, multipath_taps=0):
    """Build frames with `builder.build_frame(payload, rate)` (the reference's frame_builder in the
    tests), apply optional multipath + AWGN, concatenate with `gap` noise samples between frames.
    Returns dict(iq complex128, lts1 int64, avail int32, payloads list, rates, lengths)."""
    chunks, lts1, avail, payloads = [], [], [], []
    pos = 0
    for rate, length in zip(rates, lengths):
        pl = rng.integers(0, 256, length, dtype=np.uint8).tobytes()
        f = builder.build_frame(pl, rate)
        if multipath_taps:
            taps = (rng.standard_normal(multipath_taps) + 1j * rng.standard_normal(multipath_taps))
            taps *= np.exp(-0.7 * np.arange(multipath_taps))
            taps[0] = 1.0
            taps /= np.sqrt(np.sum(np.abs(taps) ** 2))
            f = np.convolve(f, taps)[: len(f)]
        sig = snr_sigma(f, snr_db) if snr_db is not None else 0.0
        pre = awgn(rng, gap, sig) if sig else np.zeros(gap, complex)
        body = f + (awgn(rng, len(f), sig) if sig else 0)
        chunks += [pre, body]
        lts1.append(pos + gap + LTS1_OFFSET)
        avail.append(len(f) - LTS1_OFFSET)
        payloads.append(pl)
        pos += gap + len(f)
    iq = np.concatenate(chunks)
    return dict(iq=iq, lts1=np.array(lts1, np.int64), avail=np.array(avail, np.int32), payloads=payloads,
                rates=list(rates), lengths=list(lengths))


def checker_decode(lib, corpus):
    """Run every frame of the corpus through a checker library; returns list of frame dumps."""
    out = []
    for off, n in zip(corpus["lts1"], corpus["avail"]):
        out.append(lib.decode_frame(corpus["iq"][off: off + n]))
    return out
