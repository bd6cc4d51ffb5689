"""GPU: frame detection + timing synchronisation (b200rx_sync_dev) against the reference's frame_detector and
timing_sync blocks (compiled unmodified into oracle/_ref), and the raw-samples entry points (b200rx_receive*)
against the reference's whole receiver_chain.  SURVEY 8 f1."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TAG_STS_START, TAG_STS_END, TAG_LTS1, TAG_LTS2 = 1, 2, 4, 5
DELAY = 160  # timing_sync's output is its input delayed by CARRYOVER_LENGTH (timing_sync.h:13)


def _capture(ref, rng, rates, lengths, snr_db, gap=600, lead=400, tail=8192, noise_floor=True):
    """Frames from the reference's frame_builder with gaps; AWGN over the whole capture (gaps included) so that the
    detector never sees exact zeros (see DESIGN 'Known deviations')."""
    chunks, payloads = [np.zeros(lead, complex)], []
    for rate, length in zip(rates, lengths):
        pl = rng.integers(0, 256, length, dtype=np.uint8).tobytes()
        payloads.append(pl)
        chunks += [ref.build_frame(pl, rate), np.zeros(gap, complex)]
    chunks.append(np.zeros(tail, complex))
    x = np.concatenate(chunks)
    if snr_db is not None:
        sig = np.sqrt(0.0127 / 10 ** (snr_db / 10.0) / 2.0)  # mean |x|^2 of a frame body
        x = x + sig * (rng.standard_normal(len(x)) + 1j * rng.standard_normal(len(x)))
    return x, payloads


def _gpu_sync(rx, x, phase_in=0.0):
    import torch
    dev = torch.device("cuda:0")
    d = torch.from_numpy(np.ascontiguousarray(x).view(np.float64)).to(dev)
    mf = rx.max_frames
    tags = torch.zeros(len(x), dtype=torch.uint8, device=dev)
    lts1 = torch.zeros(mf, dtype=torch.int64, device=dev)
    avail = torch.zeros(mf, dtype=torch.int32, device=dev)
    phase = torch.zeros(mf, dtype=torch.float64, device=dev)
    res = rx.sync_dev(d, phase_in, tags, lts1, avail, phase)
    nf = res["n_frames"]
    return res, tags.cpu().numpy(), lts1.cpu().numpy()[:nf], avail.cpu().numpy()[:nf], phase.cpu().numpy()[:nf]


def _ref_tags(ref, x):
    out, tags = ref.sync(x)
    # realign to the input stream: output index i carries input sample i - 160
    n = len(x)
    t = np.zeros(n, np.uint8)
    t[: n - DELAY] = tags[DELAY:n]
    rotated = np.zeros(n, complex)
    rotated[: n - DELAY] = out[DELAY:n]
    return t, rotated


@pytest.mark.parametrize("snr", [30, 20, 12])
def test_tags_match_reference_blocks(ref, rx_factory, snr):
    rng = np.random.default_rng(100 + snr)
    rates = [10, 8, 5, 3, 0, 2, 9, 7, 1, 4, 6, 10, 10, 8]
    lengths = [int(rng.integers(20, 1200)) for _ in rates]
    x, _ = _capture(ref, rng, rates, lengths, snr)
    want, rotated = _ref_tags(ref, x)
    rx = rx_factory(64, 1500)
    res, tags, lts1, avail, phase = _gpu_sync(rx, x)
    valid = len(x) - DELAY
    assert np.array_equal(tags[:valid], want[:valid]), "tag placement differs from frame_detector + timing_sync"
    ref_lts1 = np.nonzero(want == TAG_LTS1)[0]
    assert np.array_equal(lts1, ref_lts1)
    assert res["n_frames"] == len(ref_lts1) >= (10 if snr >= 20 else 1)
    # frames own the samples up to the next LTS1
    nxt = np.append(ref_lts1[1:], len(x))
    assert np.array_equal(avail.astype(np.int64), nxt - ref_lts1)
    # the constant rotation the reference applied from each frame's STS_END on (timing_sync.cpp:118-125)
    for k, p in enumerate(ref_lts1):
        i = int(p) + 200  # well inside the frame
        r = rotated[i] / x[i]
        assert abs(r - np.exp(1j * phase[k])) < 1e-9, (k, r, phase[k])
    assert res["phase_valid"] and abs(res["last_phase"] - phase[-1]) == 0.0


def test_noiseless_stream_sts_end_and_lts_tags(ref, rx_factory):
    """Exact zeros between frames: the reference divides rounding residue by rounding residue there, so STS_START
    may differ; the tags that drive the decoder (STS_END, LTS1, LTS2) must still agree."""
    rng = np.random.default_rng(7)
    rates = [10, 5, 0, 8]
    x, _ = _capture(ref, rng, rates, [300, 100, 50, 1500], None)
    want, _ = _ref_tags(ref, x)
    rx = rx_factory(16, 1500)
    res, tags, lts1, _, _ = _gpu_sync(rx, x)
    valid = len(x) - DELAY
    for t in (TAG_LTS1, TAG_LTS2):
        assert np.array_equal(np.nonzero(tags[:valid] == t)[0], np.nonzero(want[:valid] == t)[0])
    assert len(lts1) == len(rates)


@pytest.mark.parametrize("snr", [28, 16])
def test_receive_matches_reference_chain(ref, rx_factory, snr):
    """Raw capture -> payloads: b200rx_receive against receiver_chain::process_samples fed 4096-sample chunks."""
    rng = np.random.default_rng(200 + snr)
    rates = [10, 10, 8, 6, 5, 3, 2, 0, 9, 7, 4, 1, 10, 8, 5]
    lengths = [int(rng.integers(1, 1500)) for _ in rates]
    x, sent = _capture(ref, rng, rates, lengths, snr)
    chain = ref.chain_new()
    want = []
    for off in range(0, len(x), 4096):
        want += ref.chain_process(chain, x[off: off + 4096], max_len=4095, max_frames=64)
    # the reference chain is six blocks deep: flush it with silence-free noise so that the last frames come out
    sig = np.sqrt(0.0127 / 10 ** (snr / 10.0) / 2.0)
    for _ in range(8):
        pad = sig * (rng.standard_normal(4096) + 1j * rng.standard_normal(4096))
        want += ref.chain_process(chain, pad, max_len=4095, max_frames=64)
    rx = rx_factory(64, 1500)
    got, info = rx.receive(x)
    assert got == want
    assert info["n_frames"] >= len(want)
    if snr >= 25:
        assert got == sent
    # device-buffer variant gives the same frames
    import torch
    dev = torch.device("cuda:0")
    d = torch.from_numpy(np.ascontiguousarray(x).view(np.float64)).to(dev)
    mf = rx.max_frames
    payload = torch.zeros((mf, 1500), dtype=torch.uint8, device=dev)
    length = torch.zeros(mf, dtype=torch.int16, device=dev)
    rate = torch.zeros(mf, dtype=torch.uint8, device=dev)
    status = torch.full((mf,), 255, dtype=torch.uint8, device=dev)
    lts1 = torch.zeros(mf, dtype=torch.int64, device=dev)
    res = rx.receive_dev(d, payload, length, rate, status, lts1)
    rx.synchronize()
    nf = res["n_frames"]
    st = status.cpu().numpy()[:nf]
    ln = length.cpu().numpy().astype(np.uint16)[:nf]
    pl = payload.cpu().numpy()
    assert [bytes(pl[f, : ln[f]]) for f in range(nf) if st[f] == 0] == want
    assert np.array_equal(st, info["status"]) and np.array_equal(lts1.cpu().numpy()[:nf].astype(np.uint64), info["lts1"])


def test_pipelined_receive_equals_synchronous(ref, rx_factory):
    """Asynchronous b200rx_receive_dev calls rotating over pipeline lanes give what the synchronous call gives."""
    import torch
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(11)
    caps = [_capture(ref, rng, [10, 8, 3, 0, 5][: 2 + k], [900, 40, 300, 10, 1500][: 2 + k], 26)[0] for k in range(4)]
    rx = rx_factory(32, 1500)
    mf = rx.max_frames

    def outs():
        return dict(payload=torch.zeros((mf, 1500), dtype=torch.uint8, device=dev),
                    length=torch.zeros(mf, dtype=torch.int16, device=dev), rate=torch.zeros(mf, dtype=torch.uint8, device=dev),
                    status=torch.zeros(mf, dtype=torch.uint8, device=dev), lts1=torch.zeros(mf, dtype=torch.int64, device=dev),
                    n=torch.zeros(1, dtype=torch.int32, device=dev))

    d = [torch.from_numpy(np.ascontiguousarray(c).view(np.float64)).to(dev) for c in caps]
    want = []
    for k in range(4):
        o = outs()
        res = rx.receive_dev(d[k], o["payload"], o["length"], o["rate"], o["status"], o["lts1"])
        rx.synchronize()
        ref_tags, _ = _ref_tags(ref, caps[k])
        n_ref = int(np.count_nonzero(ref_tags[: len(caps[k]) - DELAY] == TAG_LTS1))
        assert res["n_frames"] == n_ref >= 2 + k  # (the reference tags the odd spurious LTS as well)
        want.append(dict({key: v.cpu().numpy().copy() for key, v in o.items()}, nf=n_ref))
        assert (want[-1]["status"][n_ref:] == 255).all() and (want[-1]["status"][:n_ref] != 255).all()
    rx.set_pipeline_depth(3)
    got = [outs() for _ in range(4)]
    for rep in range(2):
        for k in range(4):
            o = got[k]
            rx.receive_dev(d[k], o["payload"], o["length"], o["rate"], o["status"], o["lts1"], n_frames=o["n"], wait=False)
            if k == 2:
                rx.join(2)  # call 0 of this round is complete from here on (on the handle's stream)
    rx.join(0)
    rx.synchronize()
    for k in range(4):
        nf = int(got[k]["n"].item())
        assert nf == want[k]["nf"]
        for key in ("payload", "length", "rate", "status"):
            assert np.array_equal(got[k][key].cpu().numpy(), want[k][key]), (k, key)
        assert np.array_equal(got[k]["lts1"].cpu().numpy()[:nf], want[k]["lts1"][:nf])
    rx.set_pipeline_depth(1)


def test_rotation_carries_over_between_captures(ref, rx_factory):
    """phase_in: the phase a previous capture left behind rotates the samples before the first STS_END."""
    rng = np.random.default_rng(9)
    x, sent = _capture(ref, rng, [10, 8], [700, 300], 30)
    rx = rx_factory(16, 1500)
    got0, info0 = rx.receive(x, 0.0)
    got1, info1 = rx.receive(x, 1.234)
    assert got0 == got1 == sent
    assert info0["last_phase"] == info1["last_phase"]


def test_empty_and_tiny_captures(rx_factory):
    rx = rx_factory(16, 1500)
    got, info = rx.receive(np.zeros(0, complex))
    assert got == [] and info["n_frames"] == 0 and info["n_events"] == 0
    got, info = rx.receive(np.ones(100, complex))
    assert got == [] and info["n_frames"] == 0
    # a constant tone is one endless plateau: STS_START once, never an STS_END
    import torch
    dev = torch.device("cuda:0")
    d = torch.ones(2 * 5000, dtype=torch.float64, device=dev)
    tags = torch.zeros(5000, dtype=torch.uint8, device=dev)
    res = rx.sync_dev(d, 0.0, tags)
    t = tags.cpu().numpy()
    assert res["n_events"] == 0 and np.count_nonzero(t == TAG_STS_START) == 1 and np.count_nonzero(t == TAG_STS_END) == 0
    # products are 1 from sample 16 on; 15 of 16 in the window (0.9375 > 0.9) from sample 30; the run reaches 16 at 45
    assert int(np.nonzero(t == TAG_STS_START)[0][0]) == 45
