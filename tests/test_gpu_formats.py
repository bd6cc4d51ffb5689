"""GPU parity of the wire-format ingestion: fc32 / sc16 samples through the C ABI against the
checker fed the same samples widened to std::complex<double> on the host.  Widening is exact for fc32 and one
IEEE multiply for sc16, so every integer output is BIT-EXACT and the constellation is within 1e-9.

Also covers the pinned-buffer ingest of the host entry point (ingest.cu: the GPU pulls only the samples the
path uses): it must give what the device-resident path gives, for every format.
"""
import ctypes as C

import numpy as np
import pytest

from ofdm_testutil import checker_decode, make_corpus
from test_gpu_parity import compare

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def narrow(iq, fmt):
    """complex128 samples -> (array in the wire format, sc16 scale, the same samples widened back to complex128)"""
    import fun_ofdm_b200 as fo
    if fmt == fo.FMT_FC64:
        return iq, 1.0, iq
    if fmt == fo.FMT_FC32:
        w = iq.astype(np.complex64)
        return w, 1.0, w.astype(np.complex128)
    scale = float(np.max(np.abs(iq.view(np.float64)))) / 30000.0  # leave head room like an ADC would
    q = np.clip(np.rint(iq.view(np.float64) / scale), -32768, 32767).astype(np.int16)
    wide = (q.astype(np.float64) * scale).view(np.complex128)
    return q, scale, wide


def dev_decode(rx, wire, corpus, taps=True):
    dev = torch.device("cuda:0")
    n = len(corpus["lts1"])
    raw = np.ascontiguousarray(wire)
    iq = torch.from_numpy(raw.view(np.uint8)).to(dev)
    lts1 = torch.from_numpy(corpus["lts1"]).to(dev)
    avail = torch.from_numpy(corpus["avail"]).to(dev)
    payload = torch.zeros((n, rx.max_payload_bytes), dtype=torch.uint8, device=dev)
    length = torch.zeros(n, dtype=torch.int16, device=dev)
    rate = torch.zeros(n, dtype=torch.uint8, device=dev)
    status = torch.full((n,), 99, dtype=torch.uint8, device=dev)
    dbg = None
    if taps:
        max_vec = int(max((a - 128) // 80 for a in corpus["avail"])) + 1
        dbg = dict(equalized=torch.zeros((n, max_vec, 48, 2), dtype=torch.float64, device=dev),
                   decoded=torch.zeros((n, rx.max_steps // 8 + 8), dtype=torch.uint8, device=dev),
                   header_field=torch.zeros(n, dtype=torch.int32, device=dev),
                   depunct=torch.zeros((n, 2 * rx.max_steps), dtype=torch.uint8, device=dev))
    rx.decode_batch_dev(iq, lts1, avail, payload, length, rate, status, dbg)
    rx.synchronize()
    out = dict(payload=payload.cpu().numpy(), length=length.cpu().numpy().astype(np.uint16).astype(int),
               rate=rate.cpu().numpy(), status=status.cpu().numpy())
    if taps:
        out.update({k: v.cpu().numpy() for k, v in dbg.items()})
    return out


@pytest.mark.parametrize("fmt_name", ["fc32", "sc16"])
def test_wire_formats_bit_exact(ref, rx_factory, fmt_name):
    import fun_ofdm_b200 as fo
    fmt = {"fc32": fo.FMT_FC32, "sc16": fo.FMT_SC16}[fmt_name]
    rng = np.random.default_rng(404 + fmt)
    rx = rx_factory(64, 1500)
    total = 0
    try:
        # no noiseless case here: quantised noiseless BPSK puts equalised points within 1 ulp of -1, exactly on the
        # demapper's truncation boundary (qam.h:112), where the last bit of any two FFT implementations decides
        # (also true of the reference's FFTW vs the checker's shim; SURVEY 8c "parity unpinned at bit level")
        for snr in (40, 26, 17):
            rates = list(range(11)) + [10, 10, 8]
            lengths = [1500] + list(rng.integers(0, 1500, len(rates) - 1))
            corpus = make_corpus(ref, rng, rates, lengths, snr_db=snr)
            wire, scale, wide = narrow(corpus["iq"], fmt)
            rx.set_sample_format(fmt, scale)
            ref_corpus = dict(corpus, iq=wide)
            want = checker_decode(ref, ref_corpus)
            got = dev_decode(rx, wire, corpus)
            total += compare(corpus, got, want)
    finally:
        rx.set_sample_format(fo.FMT_FC64)
    assert total > 10


@pytest.mark.parametrize("fmt_name", ["fc64", "fc32", "sc16"])
def test_pinned_host_ingest_matches_device_path(ref, rx_factory, fmt_name):
    """b200rx_decode_batch on a pinned buffer (GPU pulls the useful samples over PCIe) == pageable buffer (DMA copy)
    == device-resident path, frame by frame; odd frame offsets exercise unaligned sample addresses."""
    import fun_ofdm_b200 as fo
    fmt = {"fc64": fo.FMT_FC64, "fc32": fo.FMT_FC32, "sc16": fo.FMT_SC16}[fmt_name]
    rng = np.random.default_rng(505 + fmt)
    n = 600  # > 2 * the smallest chunk (256): takes the chunked pipeline
    rx = rx_factory(640, 1500)
    rates = list(rng.integers(0, 11, n))
    lengths = list(rng.integers(0, 400, n))
    lengths[0] = 1500
    corpus = make_corpus(ref, rng, rates, lengths, snr_db=24, gap=61)
    corpus["avail"][5] = 150
    corpus["avail"][9] -= 100
    wire, scale, wide = narrow(corpus["iq"], fmt)
    lib = fo.load_library()
    try:
        rx.set_sample_format(fmt, scale)
        want = dev_decode(rx, wire, corpus, taps=False)
        pageable = rx.decode_batch(wire, corpus["lts1"], corpus["avail"])
        raw = np.ascontiguousarray(wire).view(np.uint8).reshape(-1)
        p = C.c_void_p()
        assert lib.b200rx_host_alloc(C.byref(p), raw.nbytes) == 0
        try:
            C.memmove(p, raw.ctypes.data, raw.nbytes)
            pinned_view = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(raw.nbytes,)).view(wire.dtype)
            launches0 = rx.launch_count
            pinned = rx.decode_batch(pinned_view, corpus["lts1"], corpus["avail"])
            assert rx.launch_count - launches0 > 3, "pinned buffers must take the pull path"
        finally:
            lib.b200rx_host_free(p)
    finally:
        rx.set_sample_format(fo.FMT_FC64)
    for name, got in (("pageable", pageable), ("pinned", pinned)):
        payload, length, rate, status = got
        assert np.array_equal(status, want["status"]), name
        assert np.array_equal(rate, want["rate"]), name
        assert np.array_equal(length.astype(int), want["length"]), name
        for f in range(n):
            assert bytes(payload[f, : length[f]]) == bytes(want["payload"][f, : length[f]]), (name, f)
    # and the device path itself agrees with the checker on the widened samples
    chk = checker_decode(ref, dict(corpus, iq=wide))
    assert compare(corpus, want, chk, taps=False) > n // 2


def test_receive_from_sc16_capture(ref, rx_factory):
    """Raw sc16 capture -> detector + sync + decode; same frames and payloads as the fc64 path on the widened samples."""
    import fun_ofdm_b200 as fo
    rng = np.random.default_rng(606)
    rx = rx_factory(64, 1500)
    rates = [10, 8, 5, 3, 0, 10, 9, 6]
    lengths = [300, 200, 100, 150, 40, 500, 333, 64]
    corpus = make_corpus(ref, rng, rates, lengths, snr_db=27, gap=400)
    iq = np.concatenate([corpus["iq"], np.zeros(400, complex)])
    wire, scale, wide = narrow(iq, fo.FMT_SC16)
    out_wide, info_wide = rx.receive(wide)
    try:
        rx.set_sample_format(fo.FMT_SC16, scale)
        out_q, info_q = rx.receive(wire)
    finally:
        rx.set_sample_format(fo.FMT_FC64)
    assert np.array_equal(info_q["lts1"], info_wide["lts1"])
    assert np.array_equal(info_q["status"], info_wide["status"])
    assert out_q == out_wide
    assert len(out_q) >= 6


def test_submit_wait_overlapping_calls(ref, rx_factory):
    """b200rx_submit_batch / b200rx_wait: five different batches submitted back to back (more than B200RX_MAX_INFLIGHT, so
    slots are reused) give exactly what one synchronous call per batch gives."""
    import fun_ofdm_b200 as fo
    rng = np.random.default_rng(707)
    rx = rx_factory(640, 600)
    lib = fo.load_library()
    batches = []
    for b in range(5):
        n = 560 + 10 * b  # > 2 * 256: chunked pipeline
        rates = list(rng.integers(0, 11, n))
        lengths = list(rng.integers(0, 120, n))
        batches.append(make_corpus(ref, rng, rates, lengths, snr_db=22, gap=48))
    sync = [rx.decode_batch(c["iq"], c["lts1"], c["avail"]) for c in batches]

    pinned = []

    def pin(arr):
        arr = np.ascontiguousarray(arr)
        p = C.c_void_p()
        assert lib.b200rx_host_alloc(C.byref(p), max(arr.nbytes, 1)) == 0
        C.memmove(p, arr.ctypes.data, arr.nbytes)
        pinned.append(p)
        return p

    def pin_out(nbytes):
        p = C.c_void_p()
        assert lib.b200rx_host_alloc(C.byref(p), nbytes) == 0
        C.memset(p, 0xEE, nbytes)
        pinned.append(p)
        return p

    try:
        calls = []
        for c in batches:
            n = len(c["lts1"])
            iq = pin(c["iq"])
            l = pin(c["lts1"].astype(np.uint64))
            a = pin(c["avail"].astype(np.uint32))
            o = (pin_out(n * 600), pin_out(2 * n), pin_out(n), pin_out(n))
            t = rx.submit_batch_ptr(iq, len(c["iq"]), l, a, n, o[0], 600, o[1], o[2], o[3])
            calls.append((t, n, o))
        assert len({t for t, _, _ in calls}) == 5
        rx.wait(calls[2][0])  # out of order is fine
        rx.wait(0)
        for (t, n, o), want in zip(calls, sync):
            payload = np.ctypeslib.as_array(C.cast(o[0], C.POINTER(C.c_uint8)), shape=(n, 600))
            length = np.ctypeslib.as_array(C.cast(o[1], C.POINTER(C.c_uint16)), shape=(n,))
            rate = np.ctypeslib.as_array(C.cast(o[2], C.POINTER(C.c_uint8)), shape=(n,))
            status = np.ctypeslib.as_array(C.cast(o[3], C.POINTER(C.c_uint8)), shape=(n,))
            assert np.array_equal(status, want[3]) and np.array_equal(rate, want[2]) and np.array_equal(length, want[1])
            for f in range(n):
                assert np.array_equal(payload[f, : length[f]], want[0][f, : length[f]]), f
            assert (status == 0).sum() > n // 3
    finally:
        rx.wait(0)
        for p in pinned:
            lib.b200rx_host_free(p)


@pytest.mark.parametrize("fmt_name", ["FMT_FC64", "FMT_FC32", "FMT_SC16"])
def test_two_phase_pass_in_every_sample_format(ref, rx_factory, fmt_name):
    """The pass entry points driven directly (open / put in two pieces / scan / decode / wait), graph-replayed scan and
    launch-by-launch scan, for every raw sample format: frame list, header verdicts and decoded payloads must equal what
    b200rx_receive returns for the same capture in the same format (which the other tests pin to the reference)."""
    import fun_ofdm_b200 as fo
    from fun_ofdm_b200.rx import SyncResult
    from test_gpu_sync import _capture
    fmt = getattr(fo, fmt_name)
    rng = np.random.default_rng(600 + fmt)
    rates = [int(r) for r in rng.integers(0, 11, 14)]
    lengths = [int(v) for v in rng.integers(0, 500, 14)]
    x, _ = _capture(ref, rng, rates, lengths, 40, gap=500, lead=300, tail=2048)
    wire, scale, _ = narrow(x, fmt)
    wire = np.ascontiguousarray(wire)
    rx = rx_factory(32, 600)
    rx.set_sample_format(fmt, scale)
    want_payloads, info = rx.receive(wire)
    nf = len(info["status"])
    assert nf >= 10
    L, h = rx.lib, rx.h
    bps = {fo.FMT_FC64: 16, fo.FMT_FC32: 8, fo.FMT_SC16: 4}[fmt]
    raw = wire.view(np.uint8).reshape(-1)
    n = raw.size // bps
    cut = n // 3

    class PassFrame(C.Structure):
        _fields_ = [("lts1", C.c_uint64), ("avail", C.c_uint32), ("length", C.c_uint16), ("rate", C.c_uint8), ("status", C.c_uint8),
                    ("sts_end", C.c_uint64), ("phase", C.c_double)]
    assert C.sizeof(PassFrame) == 32
    try:
        for graph in (1, 0):
            rx.set_tuning("scan_graph", graph)
            rx.set_pipeline_depth(2)
            for rep in range(3):  # lanes are reused: the second pass on a lane replays its captured graph
                frames = (PassFrame * 32)()
                res = SyncResult()
                assert L.b200rx_pass_open(h) == 0
                assert L.b200rx_pass_put(h, raw.ctypes.data, cut) == 0
                assert L.b200rx_pass_put(h, raw.ctypes.data + cut * bps, n - cut) == 0
                assert L.b200rx_pass_scan(h, 0.0, frames, 32, C.byref(res)) == 0, L.b200rx_last_error(h)
                assert res.n_frames == nf
                for f in range(nf):
                    assert frames[f].lts1 == int(info["lts1"][f])
                    hdr_ok = info["status"][f] in (fo.ST_OK, fo.ST_CRC_FAIL)
                    assert (frames[f].status == fo.ST_OK) == hdr_ok, (f, frames[f].status, info["status"][f])
                    if hdr_ok:
                        assert frames[f].length == int(info["length"][f]) and frames[f].rate == int(info["rate"][f])
                sel = (C.c_uint8 * 32)(*[1 if frames[f].status == fo.ST_OK else 0 for f in range(nf)])
                pl = np.zeros((32, 600), np.uint8)
                st = np.full(32, 200, np.uint8)
                t = C.c_uint64()
                assert L.b200rx_pass_decode(h, sel, pl.ctypes.data, 600, st.ctypes.data, C.byref(t)) == 0 and t.value != 0
                assert L.b200rx_pass_wait(h, t.value) == 0 and L.b200rx_pass_poll(h, t.value) == 1
                got = [bytes(pl[f, : frames[f].length]) for f in range(nf) if sel[f] and st[f] == fo.ST_OK]
                assert got == want_payloads, (fmt_name, graph, rep, len(got), len(want_payloads))
                for f in range(nf):
                    if sel[f]:
                        assert st[f] == info["status"][f], (f, st[f], info["status"][f])
            rx.synchronize()
            rx.set_pipeline_depth(1)
    finally:
        rx.set_tuning("scan_graph", 1)
        rx.set_sample_format(fo.FMT_FC64)
