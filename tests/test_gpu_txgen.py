"""GPU: the device-side frame generator + channel (b200tx_build_batch_dev) against the host generator
(b200tx_build_batch, itself pinned against the reference's frame_builder by tests/test_abi_and_host.py) and against
frame_builder::build_frame directly.  Coded bits are identical by construction when the samples agree to 1e-12: a
flipped coded bit moves a constellation point by >= 0.3."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

TOL = 1e-12


def _payloads(rng, lengths):
    return [rng.integers(0, 256, int(n), dtype=np.uint8).tobytes() for n in lengths]


def test_device_frames_equal_reference_frame_builder(ref):
    from fun_ofdm_b200 import tx
    rng = np.random.default_rng(11)
    rates = list(range(11)) * 2
    lengths = [0, 1, 2, 37, 100, 255, 256, 700, 1499, 1500, 4095] + [int(v) for v in rng.integers(0, 2000, 11)]
    payloads = _payloads(rng, lengths)
    got = tx.build_corpus_dev(payloads, rates, snr_db=None, lead_in=0)
    iq = got["iq"].cpu().numpy().view(np.complex128)
    for f, (pl, rate) in enumerate(zip(payloads, rates)):
        want = ref.build_frame(pl, rate)
        off = int(got["frame_off"][f])
        err = np.abs(iq[off: off + len(want)] - want).max()
        assert err < TOL, (f, rate, len(pl), err)
    assert int(got["frame_off"][-1]) + tx.frame_samples(rates[-1], lengths[-1]) == len(iq)


@pytest.mark.parametrize("taps,snr,lead", [(0, None, 0), (0, 25.0, 0), (4, 30.0, 100), (8, 22.0, 17), (0, None, 50)])
def test_device_generator_equals_host_generator(taps, snr, lead):
    from fun_ofdm_b200 import tx
    rng = np.random.default_rng(21 + taps)
    rates = [int(r) for r in rng.integers(0, 11, 40)]
    lengths = [int(v) for v in rng.integers(0, 1200, 40)]
    payloads = _payloads(rng, lengths)
    host = tx.build_corpus(payloads, rates, snr_db=snr, multipath_taps=taps, lead_in=lead, seed=77, threads=4)
    dev = tx.build_corpus_dev(payloads, rates, snr_db=snr, multipath_taps=taps, lead_in=lead, seed=77)
    a = dev["iq"].cpu().numpy().view(np.complex128)
    assert a.shape == host["iq"].shape
    assert np.array_equal(dev["lts1"].cpu().numpy(), host["lts1"].astype(np.int64))
    assert np.array_equal(dev["avail"].cpu().numpy(), host["avail"].astype(np.int32))
    err = np.abs(a - host["iq"]).max()
    assert err < TOL, err


def test_device_corpus_round_trip(rx_factory):
    """generate on the GPU -> decode on the GPU: every payload comes back (rates without noiseless self-failures)."""
    from fun_ofdm_b200 import tx
    rng = np.random.default_rng(5)
    n = 256
    rates = [[10, 9, 8, 6, 3, 0][i % 6] for i in range(n)]
    lengths = [int(v) for v in rng.integers(1, 1500, n)]
    payloads = _payloads(rng, lengths)
    c = tx.build_corpus_dev(payloads, rates, snr_db=30.0, multipath_taps=0, lead_in=0, seed=9)
    rx = rx_factory(n, 1500)
    dev = torch.device("cuda:0")
    payload = torch.zeros((n, 1500), dtype=torch.uint8, device=dev)
    length = torch.zeros(n, dtype=torch.int16, device=dev)
    rate = torch.zeros(n, dtype=torch.uint8, device=dev)
    status = torch.full((n,), 99, dtype=torch.uint8, device=dev)
    rx.decode_batch_dev(c["iq"], c["lts1"], c["avail"], payload, length, rate, status)
    rx.synchronize()
    st = status.cpu().numpy()
    pl = payload.cpu().numpy()
    ln = length.cpu().numpy().astype(np.uint16)
    assert (st == 0).sum() >= n - 2, np.bincount(st)
    for f in np.nonzero(st == 0)[0]:
        assert bytes(pl[f, : ln[f]]) == payloads[f], f
    assert np.array_equal(rate.cpu().numpy()[st == 0], np.array(rates, np.uint8)[st == 0])
