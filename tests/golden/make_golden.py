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
    print("wrote", os.path.join(HERE, "stage_kat.npz"), os.path.getsize(os.path.join(HERE, "stage_kat.npz")),
          os.path.join(HERE, "frames.npz"), os.path.getsize(os.path.join(HERE, "frames.npz")))
    print(np.array(meta))


if __name__ == "__main__":
    main()
