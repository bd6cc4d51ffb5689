"""CPU tests: the oracle port (oracle/ofdm_oracle.c) against the committed golden vectors generated from
the compiled reference (tests/golden/make_golden.py), and — where oracle/_ref exists — against the
reference itself on fresh random inputs.  Integer outputs bit-exact; fp64 outputs within 1e-12."""
import os
import sys
import zlib

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = np.load(os.path.join(HERE, "golden", "stage_kat.npz"))
FRAMES = np.load(os.path.join(HERE, "golden", "frames.npz"))
META = FRAMES["meta"]


def test_known_answers_from_survey(port):
    # SURVEY.md section 4 integer known-answers (measured on the compiled reference)
    assert list(KAT["sizeof"]) == [24, 1032, 776]
    assert int(KAT["crc_check"][0]) == 0xCBF43926 == port.crc32(b"123456789") == zlib.crc32(b"123456789")
    b0 = "".join("1" if port.parity((2 * s) & 121) else "0" for s in range(32))
    b1 = "".join("1" if port.parity((2 * s) & 91) else "0" for s in range(32))
    assert b0 == "00001111111100001111000000001111"
    assert b1 == "01011010101001010101101010100101"
    perm = port.interleave(np.arange(48, dtype=np.uint8))
    assert [int(np.nonzero(perm == k)[0][0]) for k in range(48)] == [3 * (k % 16) + k // 16 for k in range(48)]
    assert list(port.depuncture(np.array([1, 2, 3, 4], np.uint8), 2)) == [1, 2, 127, 3, 127, 4]
    assert list(port.depuncture(np.array([1, 2, 3], np.uint8), 1)) == [1, 127, 2, 3]
    q = port.demodulate(np.array([0.46 + 0.0j, -1.08 + 0j]), 10)
    assert list(q[:3]) == [223, 161, 159] and list(q[3:6]) == [128, 255, 64] and list(q[6:9]) == [0, 33, 97]


def test_stage_functions_match_golden(port):
    assert np.array_equal(port.conv_encode(KAT["enc_in"], 300), KAT["enc_out"])
    assert np.array_equal(port.interleave(KAT["ilv_in"]), KAT["ilv_out"])
    assert np.array_equal(port.deinterleave(KAT["ilv_in"]), KAT["deilv_out"])
    for rate in (1, 2):
        assert np.array_equal(port.puncture(KAT["punc_in_%d" % rate], rate), KAT["punc_out_%d" % rate])
        assert np.array_equal(port.depuncture(KAT["punc_in_%d" % rate], rate), KAT["depunc_out_%d" % rate])
    for rate in (0, 3, 6, 10):
        assert np.array_equal(port.demodulate(KAT["demod_in_%d" % rate], rate), KAT["demod_out_%d" % rate])
        assert np.array_equal(port.modulate(KAT["mod_in_%d" % rate], rate), KAT["mod_out_%d" % rate])
    assert np.abs(port.fft_forward(KAT["fft_in"]) - KAT["fft_out"]).max() < 1e-12
    for i in range(4):
        assert np.array_equal(port.conv_decode(KAT["vit_in_%d" % i], 402), KAT["vit_out_%d" % i]), i


@pytest.mark.parametrize("k", range(len(META)))
def test_frames_match_golden(port, k):
    rate, length, snr, hdr_ok, field, par, rate_valid, drate, dlen, nsym, crc_ok, nvec = (int(x) for x in META[k])
    d = port.decode_frame(FRAMES["window_%d" % k])
    assert (d.hdr_ok, d.hdr_field, d.hdr_parity, d.rate_valid) == (bool(hdr_ok), field, par, bool(rate_valid))
    assert d.n_vectors == nvec
    if not hdr_ok:
        return
    assert (d.rate, d.length, d.nsym, d.crc_ok) == (drate, dlen, nsym, bool(crc_ok))
    assert np.abs(d.eq - FRAMES["eq_%d" % k]).max() < 1e-10
    for name in ("soft", "deint", "depunct", "decoded", "descrambled"):
        assert np.array_equal(getattr(d, name), FRAMES["%s_%d" % (name, k)]), name
    if crc_ok:
        assert bytes(d.payload) == bytes(FRAMES["payload_%d" % k]) == bytes(FRAMES["tx_payload_%d" % k])


def test_port_against_reference_on_random_frames(ref, port):
    """Needs oracle/_ref (build container).  All 11 rates, clean / mild / failing SNR."""
    rng = np.random.default_rng(99)
    for rate in range(11):
        for snr in (None, 28, [3, 4, 6, 6, 8, 10, 12, 14, 16, 20, 22][rate]):
            length = int(rng.integers(0, 400))
            pl = rng.integers(0, 256, length, dtype=np.uint8).tobytes()
            f = ref.build_frame(pl, rate)
            if snr is not None:
                sig = np.sqrt(np.mean(np.abs(f[320:]) ** 2) / 10 ** (snr / 10.0) / 2.0)
                f = f + sig * (rng.standard_normal(len(f)) + 1j * rng.standard_normal(len(f)))
            a, b = ref.decode_frame(f[184:]), port.decode_frame(f[184:])
            assert (a.hdr_ok, a.hdr_field, a.crc_ok, a.rate, a.length) == (b.hdr_ok, b.hdr_field, b.crc_ok, b.rate, b.length)
            assert np.abs(a.eq - b.eq).max() < 1e-10
            if a.hdr_ok and a.n_vectors >= 1 + a.nsym:
                for name in ("soft", "depunct", "decoded", "descrambled"):
                    assert np.array_equal(getattr(a, name), getattr(b, name)), (rate, snr, name)


def test_viterbi_port_against_reference_saturating_inputs(ref, port):
    """The saturating-metric / renormalisation quirks only show on noisy inputs."""
    rng = np.random.default_rng(5)
    for sigma in (0, 30, 60, 90, 120, 200):
        for nb in (18, 90, 1002):
            d = rng.integers(0, 256, (nb + 13) // 8 + 1, dtype=np.uint8)
            coded = ref.conv_encode(d, nb).astype(np.float64) * 255.0
            soft = np.clip(np.rint(coded + sigma * rng.standard_normal(len(coded))), 0, 255).astype(np.uint8)
            assert np.array_equal(ref.conv_decode(soft, nb), port.conv_decode(soft, nb)), (sigma, nb)


def test_batch_entry_point_of_port(port):
    wins = [FRAMES["window_%d" % k] for k in range(len(META))]
    iq = np.concatenate(wins)
    off = np.cumsum([0] + [len(w) for w in wins[:-1]]).astype(np.int64)
    avail = np.array([len(w) for w in wins], np.int32)
    payload, length, status, _ = port.decode_batch(iq, off, avail, max_len=512, threads=2)
    for k in range(len(META)):
        hdr_ok, crc_ok, dlen = int(META[k][3]), int(META[k][10]), int(META[k][8])
        assert (status[k] == 0) == bool(hdr_ok and crc_ok)
        if status[k] == 0:
            assert bytes(payload[k, :dlen]) == bytes(FRAMES["payload_%d" % k])


# ---- full-size frames: every rate x {1500, 4095} bytes x {45 dB, 30 dB, the SNR where about half the frames fail} ----
BIG = np.load(os.path.join(HERE, "golden", "big_frames.npz"))


def big_frame(k):
    """Regenerates fixture frame k from its parameters (tests/golden/make_golden.py: big_frame_samples) and checks the
    SHA-256 of the samples against the one recorded when the reference decoded them.  Returns (meta row, payload, window)."""
    import hashlib
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from make_golden import big_frame_samples
    row = [int(v) for v in BIG["meta"][k]]
    rate, length, snr, seed = row[:4]
    payload, win = big_frame_samples(rate, length, None if snr < 0 else float(snr), seed)
    assert hashlib.sha256(win.tobytes()).digest() == bytes(BIG["sha256_%d" % k]), \
        "fixture %d: the regenerated samples differ from the ones the reference decoded" % k
    return row, payload, win


def test_big_fixture_covers_every_rate_and_both_verdicts():
    m = BIG["meta"]
    assert len(m) == 66 and set(m[:, 0]) == set(range(11)) and set(m[:, 1]) == {1500, 4095}
    assert m[:, 4].all()                      # every header decodes
    assert (m[:, 10] == 0).sum() >= 5         # CRC failures are in the set ...
    assert m[:, 10].sum() >= 44               # ... and so are full-size successes (all clean and 30 dB frames)


@pytest.mark.parametrize("k", range(66))
def test_port_matches_reference_on_full_size_frames(port, k):
    row, payload, win = big_frame(k)
    rate, length, snr, seed, hdr_ok, field, rate_valid, drate, dlen, nsym, crc_ok, pl_equal = row
    d = port.decode_frame(win)
    assert (d.hdr_ok, d.hdr_field, d.rate_valid) == (bool(hdr_ok), field, bool(rate_valid))
    assert (d.rate, d.length, d.nsym, d.crc_ok) == (drate, dlen, nsym, bool(crc_ok))
    assert np.array_equal(d.descrambled, BIG["descrambled_%d" % k])
    if crc_ok:
        assert (bytes(d.payload) == payload.tobytes()) == bool(pl_equal)
