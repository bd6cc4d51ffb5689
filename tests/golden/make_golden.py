#!/usr/bin/env python
"""Generates tests/golden/*.npz from the UNMODIFIED reference compiled under oracle/_ref
(`make -C oracle ref`, needs /root/reference).  Run from the repo root:

    python tests/golden/make_golden.py

The fixtures pin the oracle port (oracle/ofdm_oracle.c), the product's frame generator and the CUDA
path on machines where the reference sources are not available (the GPU box).  Frames are kept short
so the files stay small; every intermediate the reference produces on the way is stored.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bind  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


FAIL_SNR = [3, 5, 7, 6, 8, 10, 12, 14, 16, 20, 22]  # dB at which roughly half of the 1500-byte frames fail, per rate


def big_cases():
    """(rate, length, snr_db, seed) of the full-size fixtures: every rate x {1500, 4095} bytes x {45 dB, 30 dB, the SNR at
    which about half the frames fail}.  45 dB stands in for "clean": exactly noiseless frames put the constellation points
    ON the demapper's truncation boundaries (qam.h:112), where the last bit of the FFT's summation order decides a soft
    value - FFTW, the checker's shim FFT and any other correct DFT differ there (known deviations)."""
    cases = []
    for rate in range(11):
        for length in (1500, 4095):
            for k, snr in enumerate((45.0, 30.0, float(FAIL_SNR[rate]))):
                cases.append((rate, length, snr, 7000 + 100 * rate + 10 * (length == 4095) + k))
    return cases


def big_frame_samples(rate, length, snr, seed):
    """The frame as the product's host generator builds it (fun_ofdm_b200/host/txgen.cpp: counter-based noise, the same
    bits on every machine; pinned against frame_builder by the tests).  Full-size frames are not stored - 66 of them are
    33 MB of samples - but regenerated from these parameters; the fixture keeps the SHA-256 of the sample bytes, so a test
    that regenerates them knows it decodes exactly what the reference decoded when the fixture was made."""
    from fun_ofdm_b200 import tx
    payload = np.random.default_rng(seed).integers(0, 256, length, dtype=np.uint8)
    c = tx.build_corpus([payload.tobytes()], [rate], snr_db=snr, lead_in=0, seed=seed, threads=1)
    return payload, np.ascontiguousarray(c["iq"][int(c["lts1"][0]): int(c["lts1"][0]) + int(c["avail"][0])])


def big_frames(ref):
    import hashlib
    out, meta = {}, []
    for k, (rate, length, snr, seed) in enumerate(big_cases()):
        payload, win = big_frame_samples(rate, length, snr, seed)
        d = ref.decode_frame(win)
        out["sha256_%d" % k] = np.frombuffer(hashlib.sha256(win.tobytes()).digest(), dtype=np.uint8)
        out["descrambled_%d" % k] = d.descrambled if d.descrambled is not None else np.zeros(0, np.uint8)
        meta.append([rate, length, -1 if snr is None else int(round(snr)), seed, int(d.hdr_ok), d.hdr_field, int(d.rate_valid),
                     d.rate, d.length, d.nsym, int(d.crc_ok), int(d.crc_ok and bytes(d.payload) == payload.tobytes())])
    out["meta"] = np.array(meta, dtype=np.int64)
    path = os.path.join(HERE, "big_frames.npz")
    np.savez_compressed(path, **out)
    m = out["meta"]
    print("wrote", path, os.path.getsize(path), "crc ok %d of %d, header ok %d" % (m[:, 10].sum(), len(m), m[:, 4].sum()))


def main():
    ref = bind.ref()
    rng = np.random.default_rng(20261017)

    # ---- stage known-answer vectors ----
    kat = {}
    kat["sizeof"] = np.array([ref.sizeof("tagged_sample"), ref.sizeof("tagged_vector64"), ref.sizeof("tagged_vector48")])
    kat["crc_check"] = np.array([ref.crc32(b"123456789")], dtype=np.uint64)
    kat["lts_freq"] = ref.table("lts_freq")
    kat["preamble"] = ref.table("preamble")
    data = rng.integers(0, 256, 40, dtype=np.uint8)
    kat["enc_in"] = data
    kat["enc_out"] = ref.conv_encode(data, 300)
    kat["ilv_in"] = rng.integers(0, 256, 96, dtype=np.uint8)
    kat["ilv_out"] = ref.interleave(kat["ilv_in"])
    kat["deilv_out"] = ref.deinterleave(kat["ilv_in"])
    for rate in (1, 2):
        x = rng.integers(0, 256, 48, dtype=np.uint8)
        kat["punc_in_%d" % rate] = x
        kat["punc_out_%d" % rate] = ref.puncture(x, rate)
        kat["depunc_out_%d" % rate] = ref.depuncture(x, rate)
    for rate in (0, 3, 6, 10):
        pts = 1.2 * (rng.standard_normal(48) + 1j * rng.standard_normal(48))
        kat["demod_in_%d" % rate] = pts
        kat["demod_out_%d" % rate] = ref.demodulate(pts, rate)
        bits = rng.integers(0, 2, 48 * bind.RATES[rate][3], dtype=np.uint8)
        kat["mod_in_%d" % rate] = bits
        kat["mod_out_%d" % rate] = ref.modulate(bits, rate)
    x = rng.standard_normal(64) + 1j * rng.standard_normal(64)
    kat["fft_in"] = x
    kat["fft_out"] = ref.fft_forward(x)
    # Viterbi: noisy soft symbols incl. saturation-heavy cases
    for i, sigma in enumerate((0, 40, 90, 140)):
        nb = 402
        d = rng.integers(0, 256, (nb + 13) // 8 + 1, dtype=np.uint8)
        coded = ref.conv_encode(d, nb).astype(np.float64) * 255.0
        soft = np.clip(np.rint(coded + sigma * rng.standard_normal(len(coded))), 0, 255).astype(np.uint8)
        kat["vit_in_%d" % i] = soft
        kat["vit_out_%d" % i] = ref.conv_decode(soft, nb)
    np.savez_compressed(os.path.join(HERE, "stage_kat.npz"), **kat)

    # ---- whole frames: (rate, length, snr_db) ----
    cases = [(0, 60, None), (1, 45, 30), (2, 33, 30), (3, 100, 25), (4, 77, 25), (5, 64, 12), (6, 120, 25),
             (7, 90, 25), (8, 150, 18), (9, 200, 30), (10, 260, 25), (10, 0, 30), (10, 180, 17), (0, 30, 1)]
    frames = {}
    meta = []
    for k, (rate, length, snr) in enumerate(cases):
        payload = rng.integers(0, 256, length, dtype=np.uint8)
        f = ref.build_frame(payload.tobytes(), rate)
        pts = ref.ppdu_encode(payload.tobytes(), rate)
        if snr is not None:
            p = np.mean(np.abs(f[320:]) ** 2)
            sig = np.sqrt(p / 10 ** (snr / 10.0) / 2.0)
            rx = f + sig * (rng.standard_normal(len(f)) + 1j * rng.standard_normal(len(f)))
        else:
            rx = f.copy()
        win = rx[184:]
        d = ref.decode_frame(win)
        frames["tx_payload_%d" % k] = payload
        frames["tx_frame_%d" % k] = f
        frames["tx_points_%d" % k] = pts
        frames["window_%d" % k] = win
        frames["eq_%d" % k] = d.eq
        for name in ("soft", "deint", "depunct", "decoded", "descrambled", "payload"):
            v = getattr(d, name)
            frames["%s_%d" % (name, k)] = v if v is not None else np.zeros(0, np.uint8)
        meta.append([rate, length, -1 if snr is None else snr, int(d.hdr_ok), d.hdr_field, d.hdr_parity,
                     int(d.rate_valid), d.rate, d.length, d.nsym, int(d.crc_ok), d.n_vectors])
    frames["meta"] = np.array(meta, dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "frames.npz"), **frames)
    big_frames(ref)
    print("wrote", os.path.join(HERE, "stage_kat.npz"), os.path.getsize(os.path.join(HERE, "stage_kat.npz")),
          os.path.join(HERE, "frames.npz"), os.path.getsize(os.path.join(HERE, "frames.npz")))
    print(np.array(meta))


if __name__ == "__main__":
    main()
