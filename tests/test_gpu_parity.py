"""GPU parity: the CUDA path through the C ABI against the checker (reference compiled under
oracle/_ref, or the C port when that is absent) on identical samples.

Bars: payload bytes, LENGTH, rate, status, decoded (pre-descramble) bytes, header field and the
depunctured soft symbols are BIT-EXACT; equalised constellation points within 1e-9 absolute
(north star: <= 1e-4 relative to symbol energy; fp64 on both sides gives ~1e-14).
"""
import numpy as np
import pytest

from ofdm_testutil import checker_decode, make_corpus

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

EQ_TOL = 1e-9


def _checker():
    from oracle import bind
    return bind.ref() if bind.have_ref() else bind.port()


def gpu_decode(rx, corpus, taps=True):
    dev = torch.device("cuda:0")
    n = len(corpus["lts1"])
    iq = torch.from_numpy(corpus["iq"].view(np.float64)).to(dev)
    lts1 = torch.from_numpy(corpus["lts1"]).to(dev)
    avail = torch.from_numpy(corpus["avail"]).to(dev)
    stride = rx.max_payload_bytes
    payload = torch.zeros((n, max(stride, 1)), dtype=torch.uint8, device=dev)
    length = torch.zeros(n, dtype=torch.int16, device=dev)
    rate = torch.zeros(n, dtype=torch.uint8, device=dev)
    status = torch.full((n,), 99, dtype=torch.uint8, device=dev)
    dbg = None
    if taps:
        max_vec = int(max((a - 128) // 80 for a in corpus["avail"])) + 1
        dbg = dict(
            equalized=torch.zeros((n, max_vec, 48, 2), dtype=torch.float64, device=dev),
            decoded=torch.zeros((n, rx.max_steps // 8 + 8), dtype=torch.uint8, device=dev),
            header_field=torch.zeros(n, dtype=torch.int32, device=dev),
            depunct=torch.zeros((n, 2 * rx.max_steps), dtype=torch.uint8, device=dev),
        )
    rx.decode_batch_dev(iq, lts1, avail, payload, length, rate, status, dbg)
    rx.synchronize()
    out = dict(payload=payload.cpu().numpy(), length=length.cpu().numpy().astype(np.uint16).astype(int),
               rate=rate.cpu().numpy(), status=status.cpu().numpy())
    if taps:
        out.update({k: v.cpu().numpy() for k, v in dbg.items()})
    return out


def compare(corpus, got, want, taps=True):
    n = len(want)
    n_ok = 0
    for f in range(n):
        w = want[f]
        tag = "frame %d (rate %s len %s)" % (f, corpus["rates"][f], corpus["lengths"][f])
        if w.n_vectors < 1:  # not even a SIGNAL symbol in the window
            assert got["status"][f] == 4, tag
            continue
        if taps:
            assert int(got["header_field"][f]) == w.hdr_field, tag
        if not w.hdr_ok:
            assert got["status"][f] in (1, 2), tag
            assert got["status"][f] == (1 if w.hdr_parity else 2), tag
            continue
        assert got["rate"][f] == w.rate and got["length"][f] == w.length, tag
        if w.n_vectors < 1 + w.nsym:
            assert got["status"][f] == 4, tag
            continue
        assert got["status"][f] == (0 if w.crc_ok else 3), tag
        if taps:
            nv = 1 + w.nsym
            eq = got["equalized"][f, :nv].view(np.complex128)[..., 0]
            err = np.abs(eq - w.eq[:nv]).max()
            assert err < EQ_TOL, (tag, err)
            assert np.array_equal(got["depunct"][f, : len(w.depunct)], w.depunct), tag
            assert np.array_equal(got["decoded"][f, : len(w.decoded)], w.decoded), tag
        if w.crc_ok:
            n_ok += 1
            assert bytes(got["payload"][f, : w.length]) == bytes(w.payload), tag
        else:
            # payload bytes of a failing frame are still the reference's descrambled bytes
            assert bytes(got["payload"][f, : w.length]) == bytes(w.descrambled[2: 2 + w.length]), tag
    return n_ok


@pytest.mark.parametrize("rate", list(range(11)))
def test_all_rates_clean_and_noisy(ref, rx_factory, rate):
    """Config 3: every rate (8 standard + 3 non-standard 2/3), clean and at an SNR where many frames fail."""
    rng = np.random.default_rng(100 + rate)
    rx = rx_factory(64, 1500)
    fail_snr = [2, 4, 6, 5, 7, 9, 11, 13, 15, 19, 21][rate]
    total_ok = 0
    for snr in (None, 30, fail_snr):
        n = 6 if rate < 3 else 12
        lengths = [1500] + list(rng.integers(0, 1500, n - 1))
        corpus = make_corpus(ref, rng, [rate] * n, lengths, snr_db=snr)
        want = checker_decode(ref, corpus)
        got = gpu_decode(rx, corpus)
        ok = compare(corpus, got, want)
        total_ok += ok
        if snr is None and rate in (0, 3, 6, 8, 9, 10):  # rates with no noiseless self-failures
            assert ok == n
    assert total_ok > 0


def test_mixed_batch_truncated_and_garbage(ref, rx_factory):
    """Ragged input: mixed rates and lengths, truncated frames, pure noise, zero-length payload."""
    rng = np.random.default_rng(7)
    rx = rx_factory(64, 1500)
    rates = [10, 0, 5, 8, 3, 9, 10, 6, 2, 10]
    lengths = [0, 1, 37, 1500, 700, 1499, 64, 255, 100, 1000]
    corpus = make_corpus(ref, rng, rates, lengths, snr_db=28)
    corpus["avail"][2] -= 200      # truncated data part
    corpus["avail"][5] = 150       # not even a SIGNAL symbol
    corpus["avail"][7] = 208       # SIGNAL only
    # a window of pure noise
    noise_at = len(corpus["iq"])
    corpus["iq"] = np.concatenate([corpus["iq"], 0.05 * (rng.standard_normal(3000) + 1j * rng.standard_normal(3000))])
    corpus["lts1"] = np.append(corpus["lts1"], noise_at)
    corpus["avail"] = np.append(corpus["avail"], 3000).astype(np.int32)
    corpus["rates"].append(-1)
    corpus["lengths"].append(-1)
    want = checker_decode(ref, corpus)
    got = gpu_decode(rx, corpus)
    compare(corpus, got, want)
    assert got["status"][5] == 4 and got["status"][7] == 4 and got["status"][2] == 4


def test_host_buffer_entry_point_matches_device_entry_point(ref, rx_factory):
    rng = np.random.default_rng(11)
    rx = rx_factory(64, 1500)
    corpus = make_corpus(ref, rng, [10, 8, 0, 4], [1500, 300, 20, 999], snr_db=26)
    got_dev = gpu_decode(rx, corpus, taps=False)
    payload, length, rate, status = rx.decode_batch(corpus["iq"], corpus["lts1"], corpus["avail"])
    assert np.array_equal(status, got_dev["status"])
    assert np.array_equal(length.astype(int), got_dev["length"])
    assert np.array_equal(rate, got_dev["rate"])
    assert np.array_equal(payload, got_dev["payload"])
    want = checker_decode(ref, corpus)
    for f, w in enumerate(want):
        if w.crc_ok:
            assert bytes(payload[f, : w.length]) == bytes(w.payload)


def test_multipath(ref, rx_factory):
    """Config 5 flavour: <= 8-tap multipath + AWGN, genie tags; verdicts and bytes identical."""
    rng = np.random.default_rng(23)
    rx = rx_factory(64, 4095)
    rates = list(rng.choice([0, 2, 3, 5, 6, 8, 9, 10], 24))
    lengths = list(rng.integers(64, 4096, 24))
    corpus = make_corpus(ref, rng, rates, lengths, snr_db=30, multipath_taps=4)
    want = checker_decode(ref, corpus)
    got = gpu_decode(rx, corpus)
    compare(corpus, got, want)


def test_viterbi_only_entry_point(ref, rx_factory):
    """Config 4: conv_encode -> puncture -> noise -> depuncture -> b200rx_viterbi_batch_dev vs viterbi::conv_decode."""
    rng = np.random.default_rng(5)
    rx = rx_factory(64, 1500)
    dev = torch.device("cuda:0")
    cases = []
    for rate, sigma in [(0, 0), (0, 60), (1, 40), (2, 10), (10, 60), (10, 90), (9, 90), (8, 120)]:
        for nbits in (18, 90, 2042, 12090):
            steps = nbits + 6
            if rate in (2, 5, 8, 10):
                steps -= steps % 6          # whole puncture periods (3 steps)
            elif rate in (1, 4, 7, 9):
                steps -= steps % 2
            nb = steps - 6
            data = rng.integers(0, 256, (nb + 6 + 7) // 8 + 1, dtype=np.uint8)
            coded = ref.conv_encode(data, nb)
            tx = ref.puncture(coded, rate).astype(np.float64) * 255.0
            rxs = np.clip(np.rint(tx + sigma * rng.standard_normal(len(tx))), 0, 255).astype(np.uint8)
            dep = ref.depuncture(rxs, rate)
            assert len(dep) == 2 * steps
            cases.append((nb, dep, ref.conv_decode(dep, nb)))
    n = len(cases)
    stride = 2 * rx.max_steps
    sym = np.zeros((n, stride), np.uint8)
    for i, (nb, dep, _) in enumerate(cases):
        sym[i, : len(dep)] = dep
    bits = np.array([c[0] for c in cases], np.int32)
    out = torch.zeros((n, rx.max_steps // 8 + 8), dtype=torch.uint8, device=dev)
    rx.viterbi_batch_dev(torch.from_numpy(sym).to(dev), torch.from_numpy(bits).to(dev), int(bits.max()), out)
    rx.synchronize()
    out = out.cpu().numpy()
    for i, (nb, _, want) in enumerate(cases):
        assert np.array_equal(out[i, : len(want)], want), (i, nb)


def test_golden_frames_on_gpu(rx_factory):
    """Committed fixtures generated from the compiled reference (tests/golden/make_golden.py): does not
    need oracle/_ref at run time."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frames.npz"))
    meta = g["meta"]
    wins = [g["window_%d" % k] for k in range(len(meta))]
    corpus = dict(iq=np.concatenate(wins),
                  lts1=np.cumsum([0] + [len(w) for w in wins[:-1]]).astype(np.int64),
                  avail=np.array([len(w) for w in wins], np.int32))
    rx = rx_factory(64, 1500)
    got = gpu_decode(rx, corpus)
    for k in range(len(meta)):
        rate, length, snr, hdr_ok, field, par, rate_valid, drate, dlen, nsym, crc_ok, nvec = (int(x) for x in meta[k])
        assert int(got["header_field"][k]) == field
        if not hdr_ok:
            assert got["status"][k] == (1 if par else 2)
            continue
        assert (got["rate"][k], got["length"][k]) == (drate, dlen)
        assert got["status"][k] == (0 if crc_ok else 3)
        eq = got["equalized"][k, : 1 + nsym].view(np.complex128)[..., 0]
        assert np.abs(eq - g["eq_%d" % k][: 1 + nsym]).max() < EQ_TOL
        dep = g["depunct_%d" % k]
        assert np.array_equal(got["depunct"][k, : len(dep)], dep)
        dec = g["decoded_%d" % k]
        assert np.array_equal(got["decoded"][k, : len(dec)], dec)
        if crc_ok:
            assert bytes(got["payload"][k, :dlen]) == bytes(g["payload_%d" % k])


def test_config2_full_batch_properties_and_sampled_parity(rx_factory):
    """BASELINE config 2 at full size (4096 x 1500 B, 54 Mbps, 25 dB) from the product's generator:
    size-independent properties on every frame (a CRC-OK frame carries exactly the transmitted payload;
    decoding twice gives identical output), and bit-exact parity with the checker on a 192-frame sample."""
    from fun_ofdm_b200 import tx
    rng = np.random.default_rng(0xB200)
    n = 4096
    payloads = rng.integers(0, 256, (n, 1500), dtype=np.uint8)
    corpus = tx.build_corpus(payloads, np.full(n, 10, np.uint8), snr_db=25.0, seed=0xB200)
    c = dict(iq=corpus["iq"], lts1=corpus["lts1"].astype(np.int64), avail=corpus["avail"].astype(np.int32))
    rx = rx_factory(n, 1500)
    a = gpu_decode(rx, c, taps=False)
    b = gpu_decode(rx, c, taps=False)
    for key in ("payload", "length", "rate", "status"):
        assert np.array_equal(a[key], b[key]), key
    ok = a["status"] == 0
    assert ok.sum() > 0.9 * n
    assert set(np.unique(a["status"])) <= {0, 3}
    assert np.all(a["length"] == 1500) and np.all(a["rate"] == 10)
    assert np.array_equal(a["payload"][ok], payloads[ok])
    # the host-buffer entry point (chunked H2D / decode / D2H pipeline above 512 frames) must agree
    hp, hl, hr, hs = rx.decode_batch(c["iq"], c["lts1"], c["avail"])
    assert np.array_equal(hs, a["status"]) and np.array_equal(hp, a["payload"])
    assert np.array_equal(hl.astype(int), a["length"]) and np.array_equal(hr, a["rate"])
    checker = _checker()
    idx = np.concatenate([np.arange(96), np.nonzero(~ok)[0][:32], rng.integers(0, n, 64)])
    for f in idx:
        off, m = int(c["lts1"][f]), int(c["avail"][f])
        w = checker.decode_frame(c["iq"][off: off + m])
        assert (a["status"][f] == 0) == w.crc_ok, f
        want = w.payload if w.crc_ok else w.descrambled[2: 2 + w.length]
        assert bytes(a["payload"][f, : w.length]) == bytes(want), f


def test_worst_case_length(rx_factory):
    """Longest frames the 12-bit LENGTH field allows (4095 B): 32 832 trellis steps at BPSK 1/2."""
    from fun_ofdm_b200 import tx
    rng = np.random.default_rng(3)
    rates = [0, 10, 2, 9]
    payloads = [rng.integers(0, 256, 4095, dtype=np.uint8).tobytes() for _ in rates]
    corpus = tx.build_corpus(payloads, rates, snr_db=30.0, seed=1)
    c = dict(iq=corpus["iq"], lts1=corpus["lts1"].astype(np.int64), avail=corpus["avail"].astype(np.int32),
             rates=rates, lengths=[4095] * 4)
    rx = rx_factory(8, 4095)
    got = gpu_decode(rx, c)
    want = checker_decode(_checker(), c)
    assert compare(c, got, want) >= 2
    # a handle sized for short payloads must refuse them, not overflow
    small = rx_factory(8, 100)
    g2 = gpu_decode(small, c, taps=False)
    assert np.all(g2["status"] == 5)


def test_pipelined_calls_match_sequential(ref, rx_factory):
    """b200rx_set_pipeline_depth(3): consecutive calls overlap on three lanes; every call must give exactly what
    it gives alone, whatever runs next to it."""
    rng = np.random.default_rng(77)
    rx = rx_factory(64, 1500)
    dev = torch.device("cuda:0")
    corpora = []
    for k in range(7):
        n = int(rng.integers(5, 40))
        rates = list(rng.integers(0, 11, n))
        lengths = list(rng.integers(0, 1500, n))
        corpora.append(make_corpus(ref, rng, rates, lengths, snr_db=[None, 30, 12][k % 3]))
    seq = [gpu_decode(rx, c, taps=False) for c in corpora]
    rx.set_pipeline_depth(3)
    bufs = []
    for c in corpora:
        n = len(c["lts1"])
        iq = torch.from_numpy(c["iq"].view(np.float64)).to(dev)
        l = torch.from_numpy(c["lts1"]).to(dev)
        a = torch.from_numpy(c["avail"]).to(dev)
        o = (torch.zeros((n, 1500), dtype=torch.uint8, device=dev), torch.zeros(n, dtype=torch.int16, device=dev),
             torch.zeros(n, dtype=torch.uint8, device=dev), torch.full((n,), 99, dtype=torch.uint8, device=dev))
        bufs.append((iq, l, a, o))
    for iq, l, a, o in bufs:            # all seven issued without waiting
        rx.decode_batch_dev(iq, l, a, *o)
    rx.join(0)
    rx.synchronize()
    for (iq, l, a, o), s in zip(bufs, seq):
        assert np.array_equal(o[3].cpu().numpy(), s["status"])
        assert np.array_equal(o[0].cpu().numpy(), s["payload"])
        assert np.array_equal(o[1].cpu().numpy().astype(np.uint16).astype(int), s["length"])
        assert np.array_equal(o[2].cpu().numpy(), s["rate"])
    rx.set_pipeline_depth(1)
    again = gpu_decode(rx, corpora[0], taps=False)
    assert np.array_equal(again["payload"], seq[0]["payload"]) and np.array_equal(again["status"], seq[0]["status"])


def test_full_size_fixture_frames(rx_factory):
    """The committed full-size fixtures (tests/golden/big_frames.npz: all 11 rates x {1500, 4095} bytes x {45 dB, 30 dB,
    failing SNR}, decoded by the compiled reference when the fixture was made): status, rate, LENGTH, the descrambled
    bytes of every frame (CRC failures included) and the payloads must be identical.  Needs no reference on the box."""
    from test_oracle_golden import BIG, big_frame
    rx = rx_factory(66, 4095)
    rows, wins = [], []
    for k in range(66):
        row, payload, win = big_frame(k)
        rows.append((row, payload))
        wins.append(win)
    lts1 = np.cumsum([0] + [len(w) for w in wins[:-1]]).astype(np.uint64)
    corpus = dict(iq=np.concatenate(wins), lts1=lts1, avail=np.array([len(w) for w in wins], np.uint32))
    dev = torch.device("cuda:0")
    n = 66
    payload_t = torch.zeros((n, 4095), dtype=torch.uint8, device=dev)
    length = torch.zeros(n, dtype=torch.int16, device=dev)
    rate = torch.zeros(n, dtype=torch.uint8, device=dev)
    status = torch.full((n,), 99, dtype=torch.uint8, device=dev)
    decoded = torch.zeros((n, rx.max_steps // 8 + 8), dtype=torch.uint8, device=dev)
    rx.decode_batch_dev(torch.from_numpy(corpus["iq"].view(np.float64)).to(dev), torch.from_numpy(corpus["lts1"]).to(dev),
                        torch.from_numpy(corpus["avail"]).to(dev), payload_t, length, rate, status, dict(decoded=decoded))
    rx.synchronize()
    st, ln, rt, pl = status.cpu().numpy(), length.cpu().numpy().astype(np.uint16), rate.cpu().numpy(), payload_t.cpu().numpy()
    # the descrambler (ppdu.cpp:255-264) on the tapped Viterbi output, to compare failed frames byte for byte
    scr, state = [], 93
    for _ in range(127):
        fb = ((state >> 6) & 1) ^ ((state >> 3) & 1)
        scr.append(fb)
        state = ((state << 1) & 0x7E) | fb
    scr = np.array(scr, np.uint8)
    dec = decoded.cpu().numpy()
    for k, (row, payload) in enumerate(rows):
        r, le, snr, seed, hdr_ok, field, rate_valid, drate, dlen, nsym, crc_ok, pl_equal = row
        assert hdr_ok and st[k] == (0 if crc_ok else 3), (k, row, st[k])
        assert rt[k] == drate and ln[k] == dlen, (k, row)
        want = BIG["descrambled_%d" % k]
        got = dec[k, : len(want)] ^ scr[np.arange(len(want)) % 127]
        assert np.array_equal(got, want), (k, row)
        if crc_ok:
            assert (bytes(pl[k, :dlen]) == payload.tobytes()) == bool(pl_equal), (k, row)


def test_copy_counters_is_ordered_behind_its_own_call(ref, rx_factory):
    """b200rx_copy_counters: with calls pipelined over lanes every call's four counters land where they were asked for,
    although the lane's next batch resets the live counters (the race a copy on another stream would lose)."""
    rng = np.random.default_rng(21)
    rx = rx_factory(64, 600)
    corpora = [make_corpus(ref, rng, [10, 3, 0, 8][: 2 + k % 3], [500, 77, 10, 300][: 2 + k % 3], snr_db=28) for k in range(7)]
    dev = torch.device("cuda:0")
    rx.set_pipeline_depth(3)
    rows = torch.zeros((len(corpora), 4), dtype=torch.int64, device=dev)
    keep = []
    for k, c in enumerate(corpora):
        n = len(c["lts1"])
        t = [torch.from_numpy(c["iq"].view(np.float64)).to(dev), torch.from_numpy(c["lts1"]).to(dev), torch.from_numpy(c["avail"]).to(dev),
             torch.zeros((n, 600), dtype=torch.uint8, device=dev), torch.zeros(n, dtype=torch.int16, device=dev),
             torch.zeros(n, dtype=torch.uint8, device=dev), torch.zeros(n, dtype=torch.uint8, device=dev)]
        keep.append(t)
        torch.cuda.synchronize()
        rx.decode_batch_dev(*t)
        rx.copy_counters(rows[k].data_ptr())
    rx.join(0)
    rx.synchronize()
    rx.set_pipeline_depth(1)
    got = rows.cpu().numpy()
    for k, c in enumerate(corpora):
        st = keep[k][6].cpu().numpy()
        ln = keep[k][4].cpu().numpy().astype(np.uint16).astype(np.int64)
        assert got[k, 0] == int((st == 0).sum()) and got[k, 0] + got[k, 1] == len(st), (k, got[k], st)
        assert got[k, 2] == int(ln[st == 0].sum()), (k, got[k])


def test_kernel_variants_are_bit_identical(ref, rx_factory):
    """include/b200rx.h promises that b200rx_set_tuning never changes results.  Every variant of the ACS kernel
    (generation 2 / 3, lanes per frame, warps per CTA, eager / lazy renormalisation) and of the front end (split /
    one kernel) decodes the same mixed batch - all rates, ragged lengths, an SNR at which a third of the frames fail -
    to the same status, LENGTH, rate, payload bytes, Viterbi output and depunctured soft symbols as the default, and the
    default equals the reference."""
    rng = np.random.default_rng(1234)
    n = 96
    rates = [int(r) for r in rng.integers(0, 11, n)]
    lengths = [int(v) for v in rng.integers(0, 700, n)]
    corpus = make_corpus(ref, rng, rates, lengths, snr_db=14)
    rx = rx_factory(n, 1500)
    base = gpu_decode(rx, corpus)
    compare(corpus, base, checker_decode(ref, corpus))
    assert 10 < int((base["status"] == 3).sum()) < n - 10   # CRC failures and successes both present
    variants = [dict(acs_gen=3, acs_lb=2), dict(acs_gen=3, acs_lb=3), dict(acs_gen=3, acs_rn=0), dict(acs_gen=3, acs_warps=1),
                dict(acs_gen=3, acs_warps=2), dict(acs_gen=3, fe_split=0),
                dict(acs_gen=2), dict(acs_gen=2, acs_lb=2), dict(acs_gen=2, acs_lb=3), dict(acs_gen=2, acs_lb=4),
                dict(acs_gen=2, acs_lb=5), dict(acs_gen=2, acs_rn=0), dict(acs_gen=2, acs_warps=1), dict(acs_gen=2, acs_warps=4)]
    defaults = dict(acs_gen=3, acs_lb=0, acs_warps=0, acs_rn=1, fe_split=1)
    try:
        for v in variants:
            for key, val in dict(defaults, **v).items():
                rx.set_tuning(key, val)
            got = gpu_decode(rx, corpus)
            for name in ("status", "length", "rate", "payload", "decoded", "header_field", "depunct"):
                assert np.array_equal(got[name], base[name]), (v, name)
            assert np.array_equal(got["equalized"], base["equalized"]) or np.abs(got["equalized"] - base["equalized"]).max() < 1e-12, v
    finally:
        for key, val in defaults.items():
            rx.set_tuning(key, val)
