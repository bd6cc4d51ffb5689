"""GPU: fun::b200_receiver_chain (raw samples in, payloads out, chunk by chunk - SURVEY 8 f2) against the reference's
own receiver_chain::process_samples (receiver_chain.cpp:106-126, compiled unmodified into oracle/_ref) on the same
chunked stream.  Payload SEQUENCES must be identical; the GPU chain delivers a frame in the call that brings its last
sample, the reference up to five calls later, so per-call alignment is not compared."""
import ctypes as C
import os

import numpy as np
import pytest

from test_gpu_sync import _capture

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Chain:
    def __init__(self, max_frames=256, max_payload=4095, lib_path=None, depth=None, max_lag=5):
        self.lib = C.CDLL(lib_path or os.path.join(ROOT, "fun_ofdm_b200", "lib", "libb200host.so"))
        self.lib.b200host_chain_new.restype = C.c_void_p
        self.lib.b200host_chain_new.argtypes = [C.c_int, C.c_uint, C.c_uint]
        self.lib.b200host_chain_new2.restype = C.c_void_p
        self.lib.b200host_chain_new2.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_uint, C.c_uint]
        self.lib.b200host_chain_run.restype = C.c_int
        self.lib.b200host_chain_run.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_long, C.c_int, C.c_void_p, C.c_int,
                                                C.c_void_p, C.c_int, C.c_void_p]
        self.lib.b200host_alloc_samples.restype = C.c_void_p
        self.lib.b200host_alloc_samples.argtypes = [C.c_long]
        self.lib.b200host_free_samples.argtypes = [C.c_void_p]
        self.lib.b200host_chain_process.restype = C.c_int
        self.lib.b200host_chain_process.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        self.lib.b200host_chain_delete.argtypes = [C.c_void_p]
        self.lib.b200host_chain_counters.argtypes = [C.c_void_p, C.c_void_p]
        if depth is None:
            self.h = self.lib.b200host_chain_new(0, max_frames, max_payload)
        else:
            self.h = self.lib.b200host_chain_new2(0, max_frames, max_payload, depth, max_lag)
        assert self.h, "b200_receiver_chain could not be created (no GPU?)"

    def run(self, samples, chunk, by_value=False, max_out=4096, stride=4095, pinned=False):
        """The whole feed loop + flush in native code (host_capi.cpp: b200host_chain_run).  Returns (payloads, seconds of
        the feed loop, seconds including the flush).  pinned: the stream is first copied into pinned memory (not timed)."""
        payload = np.zeros((max_out, stride), np.uint8)
        length = np.zeros(max_out, np.int32)
        sec = np.zeros(2, np.float64)
        iq = np.ascontiguousarray(samples, dtype=np.complex128).view(np.float64)
        ptr, pin = iq.ctypes.data, None
        if pinned:
            pin = self.lib.b200host_alloc_samples(len(iq) // 2)
            assert pin
            C.memmove(pin, iq.ctypes.data, iq.nbytes)
            ptr = pin
        n = self.lib.b200host_chain_run(self.h, ptr, len(iq) // 2, chunk, 1 if by_value else 0, payload.ctypes.data, stride,
                                        length.ctypes.data, max_out, sec.ctypes.data)
        if pin:
            self.lib.b200host_free_samples(pin)
        assert n <= max_out
        return [bytes(payload[i, : length[i]]) for i in range(n)], float(sec[0]), float(sec[1])

    def process(self, samples, max_out=512, stride=4095):
        """samples=None: flush"""
        payload = np.zeros((max_out, stride), np.uint8)
        length = np.zeros(max_out, np.int32)
        if samples is None:
            n = self.lib.b200host_chain_process(self.h, None, -1, payload.ctypes.data, stride, length.ctypes.data, max_out)
        else:
            iq = np.ascontiguousarray(samples, dtype=np.complex128).view(np.float64)
            n = self.lib.b200host_chain_process(self.h, iq.ctypes.data, len(iq) // 2, payload.ctypes.data, stride,
                                                length.ctypes.data, max_out)
        assert n <= max_out
        return [bytes(payload[i, : length[i]]) for i in range(n)]

    def set_tuning(self, key, value):
        self.lib.b200host_chain_set_tuning.restype = C.c_int
        self.lib.b200host_chain_set_tuning.argtypes = [C.c_void_p, C.c_char_p, C.c_longlong]
        assert self.lib.b200host_chain_set_tuning(self.h, key.encode(), int(value)) == 0

    def counters(self):
        c = np.zeros(7, np.uint64)
        self.lib.b200host_chain_counters(self.h, c.ctypes.data)
        return dict(zip(["samples", "calls", "found", "ok", "crc_fail", "headers_bad", "truncated"], (int(x) for x in c)))

    def close(self):
        self.lib.b200host_chain_delete(self.h)


def _reference_chain(ref, x, chunk):
    chain = ref.chain_new()
    out = []
    for s in range(0, len(x), chunk):
        out += ref.chain_process(chain, x[s: s + chunk])
    for _ in range(8):  # drain the six-stage pipeline (one round per block)
        out += ref.chain_process(chain, np.zeros(chunk, complex))
    return out


@pytest.mark.parametrize("chunk", [4096, 1000, 333])
@pytest.mark.parametrize("snr", [28, 16])
def test_chunked_stream_matches_reference_chain(ref, chunk, snr):
    rng = np.random.default_rng(900 + chunk + snr)
    rates = [10, 8, 0, 5, 10, 3, 9, 6, 10, 2, 1, 4, 7, 10]
    lengths = [1500, 300, 40, 700, 64, 1000, 1499, 255, 0, 120, 77, 410, 900, 333]
    x, payloads = _capture(ref, rng, rates, lengths, snr, gap=500, lead=300, tail=4096)
    want = _reference_chain(ref, x, chunk)
    ch = Chain()
    got = []
    for s in range(0, len(x), chunk):
        got += ch.process(x[s: s + chunk])
    got += ch.process(None)
    c = ch.counters()
    ch.close()
    assert got == want, (len(got), len(want), c)
    assert len(got) >= (len(rates) - 3 if snr >= 25 else 3), c
    assert c["samples"] >= len(x)


def test_loopback_config1(ref):
    """BASELINE config 1 = examples/test_sim.cpp:43-104: RATE_3_4_QAM16, the 100-character string x 15 = 1500 bytes, 100
    identical frames concatenated back to back, noiseless, a zero pad behind them, fed in 4096-sample chunks.  The
    reference chain and the GPU chain must return the same payload sequence (test_sim prints `Received N packets`)."""
    data = b"I'm a little tea pot, short and stout.....here is my handle.....blah blah blah.....this rhyme sucks!"
    assert len(data) == 100
    payload = data * 15
    frame = ref.build_frame(payload, 8)
    x = np.concatenate([frame] * 100 + [np.zeros(10 * len(frame), complex)])
    want = _reference_chain(ref, x, 4096)
    ch = Chain()
    got = []
    for s in range(0, len(x), 4096):
        got += ch.process(x[s: s + 4096])
    got += ch.process(None)
    c = ch.counters()
    ch.close()
    assert got == want, (len(got), len(want), c)
    assert len(got) >= 99 and all(g == payload for g in got)


def test_one_long_call_equals_many_short_ones(ref):
    rng = np.random.default_rng(31)
    rates = [int(r) for r in rng.integers(0, 11, 40)]
    lengths = [int(v) for v in rng.integers(0, 600, 40)]
    x, _ = _capture(ref, rng, rates, lengths, 24, gap=420, lead=300, tail=2048)
    a = Chain()
    one = a.process(x) + a.process(None)
    a.close()
    b = Chain()
    many = []
    for s in range(0, len(x), 2500):
        many += b.process(x[s: s + 2500])
    many += b.process(None)
    b.close()
    assert one == many and len(one) >= 30


def test_randomised_streams_match_reference_chain(ref):
    """tools/stress_receive.py in small: random gaps (including none), rates, lengths, SNRs, frames cut short, scaled,
    overlapping, loud noise bursts; the reference chain, one GPU capture and the chunked GPU chain must agree."""
    import fun_ofdm_b200 as fo
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from stress_receive import make_stream
    rng = np.random.default_rng(4711)
    rx = fo.Receiver(0, 256, 4095)
    total = 0
    for s in range(16):
        x, snr, nf = make_stream(ref, rng)
        want1 = _reference_chain(ref, x, len(x))          # the whole stream in one call
        got1, _ = rx.receive(np.concatenate([x, np.zeros(400, complex)]))
        chunk = int(rng.choice([333, 1000, 4096, 20000]))
        want2 = _reference_chain(ref, x, chunk)           # the reference's answer depends on the chunking (timing_sync.cpp:102)
        ch = Chain(max_frames=256)
        got2 = []
        for pos in range(0, len(x), chunk):
            got2 += ch.process(x[pos: pos + chunk])
        for _ in range(8):
            got2 += ch.process(np.zeros(chunk, complex))
        ch.close()
        assert got1 == want1, (s, len(got1), len(want1), snr)
        assert got2 == want2, (s, chunk, len(got2), len(want2), snr)
        total += len(want2)
    rx.close()
    assert total > 40


@pytest.mark.parametrize("depth,lag", [(1, 0), (2, 1), (6, 5), (12, 11)])
def test_pipelined_passes_deliver_the_same_sequence(ref, depth, lag):
    """The two-phase passes (b200rx_pass_*): whatever the number of passes in flight and the delivery lag, the payload
    SEQUENCE equals the reference chain's; only the call a frame surfaces in moves (the reference itself delivers up to
    five calls late, receiver_chain.cpp:118-125)."""
    rng = np.random.default_rng(77 + depth)
    rates = [int(r) for r in rng.integers(0, 11, 24)]
    lengths = [int(v) for v in rng.integers(0, 900, 24)]
    x, _ = _capture(ref, rng, rates, lengths, 22, gap=450, lead=300, tail=4096)
    for chunk in (4096, 1500):
        want = _reference_chain(ref, x, chunk)
        ch = Chain(depth=depth, max_lag=lag)
        got, lagged = [], 0
        for s in range(0, len(x), chunk):
            out = ch.process(x[s: s + chunk])
            got += out
        tail = ch.process(None)
        lagged = len(tail)
        got += tail
        c = ch.counters()
        ch.close()
        assert got == want, (depth, lag, chunk, len(got), len(want), c)
        assert len(got) >= 12, c
        if lag == 0:
            assert lagged <= 1  # synchronous: nothing but a frame completed by the flush pad is left for the flush


def test_native_feed_loop_by_value_pointer_and_pinned(ref):
    """b200host_chain_run = the loop of examples/test_sim.cpp:77-97 in native code: by-value std::vector like test_sim,
    the pointer overload, and a pinned caller buffer with calls long enough (>= 65 536 samples) to be copied to the GPU
    straight from it - all against the reference chain fed the same chunks."""
    rng = np.random.default_rng(2027)
    rates = [int(r) for r in rng.integers(0, 11, 60)]
    lengths = [int(v) for v in rng.integers(0, 1200, 60)]
    x, _ = _capture(ref, rng, rates, lengths, 24, gap=420, lead=300, tail=8192)
    for chunk, by_value, pinned in ((4096, True, False), (4096, False, False), (70000, False, True), (70000, True, False),
                                    (200000, False, True)):
        want = _reference_chain(ref, x, chunk)
        ch = Chain(max_frames=512)
        got, t_feed, t_all = ch.run(x, chunk, by_value=by_value, pinned=pinned)
        c = ch.counters()
        ch.close()
        assert got == want, (chunk, by_value, pinned, len(got), len(want), c)
        assert len(got) >= 40 and t_all >= t_feed > 0.0


def test_receiver_wrapper_callback_pause_resume(ref):
    """fun::b200_receiver = the reference's fun::receiver loop (receiver.cpp:42-58: get_samples -> process_samples ->
    callback, pause / resume around it) over a sample-source functor: same payload sequence as the reference chain fed
    the same 4096-sample rounds, and no round runs while the receiver is paused."""
    rng = np.random.default_rng(515)
    rates = [int(r) for r in rng.integers(0, 11, 30)]
    lengths = [int(v) for v in rng.integers(0, 800, 30)]
    x, _ = _capture(ref, rng, rates, lengths, 25, gap=450, lead=300, tail=6 * 4096)
    x = x[: (len(x) // 4096) * 4096]
    want = _reference_chain(ref, x, 4096)
    lib = C.CDLL(os.path.join(ROOT, "fun_ofdm_b200", "lib", "libb200host.so"))
    lib.b200host_receiver_run.restype = C.c_int
    lib.b200host_receiver_run.argtypes = [C.c_void_p, C.c_long, C.c_long, C.c_long, C.c_int, C.c_uint, C.c_uint, C.c_void_p, C.c_int,
                                          C.c_void_p, C.c_int, C.c_void_p]
    iq = np.ascontiguousarray(x).view(np.float64)
    pl = np.zeros((256, 4095), np.uint8)
    ln = np.zeros(256, np.int32)
    paused = C.c_long(-5)
    n = lib.b200host_receiver_run(iq.ctypes.data, len(x), 4096, 5, 40, 256, 4095, pl.ctypes.data, 4095, ln.ctypes.data, 256,
                                  C.byref(paused))
    assert n >= 20
    got = [bytes(pl[i, : ln[i]]) for i in range(n)]
    assert got == want, (len(got), len(want))
    assert paused.value == 0


def test_flush_starts_a_new_stream(ref):
    """After flush() the chain behaves like a fresh one: a frame that starts within the first samples of the next stream
    is found (a retained buffer's first 256 samples are skipped only when the stream did not start there)."""
    data = bytes(range(200)) * 3
    frame = ref.build_frame(data, 8)
    x = np.concatenate([np.tile(frame, 6), np.zeros(3 * len(frame), complex)])
    want = _reference_chain(ref, x, 4096)
    ch = Chain()
    first, _, _ = ch.run(x, 4096)
    second, _, _ = ch.run(x, 4096)
    ch.close()
    assert first == want and second == want and len(want) == 6


def test_scan_as_graph_and_launch_by_launch_agree(ref):
    """The scan phase of a pass replayed as a CUDA graph (kernels read the per-call scalars from the lane's parameter
    block) and issued launch by launch (scalars in the kernel arguments) are the same computation: identical payload
    sequences, equal to the reference chain's, for small and large calls and across staging growth."""
    rng = np.random.default_rng(4242)
    rates = [int(r) for r in rng.integers(0, 11, 36)]
    lengths = [int(v) for v in rng.integers(0, 1000, 36)]
    x, _ = _capture(ref, rng, rates, lengths, 23, gap=430, lead=300, tail=8192)
    for chunk in (4096, 333, 50000):
        want = _reference_chain(ref, x, chunk)
        outs = []
        for graph in (1, 0):
            ch = Chain(max_frames=256)
            ch.set_tuning("scan_graph", graph)
            got, _, _ = ch.run(x, chunk)
            ch.close()
            outs.append(got)
        assert outs[0] == outs[1] == want, (chunk, len(outs[0]), len(outs[1]), len(want))
        assert len(want) >= 20


def test_pass_api_rejects_misuse_and_releases_memory(ref):
    """Order and format errors of the two-phase pass entry points come back as B200RX_E_ARG (never a crash, never a
    silent fallback), and handles / chains give their device memory back."""
    import torch
    import fun_ofdm_b200 as fo
    from fun_ofdm_b200.rx import SyncResult
    E_ARG = -1
    rx = fo.Receiver(0, 32, 600)
    L, h = rx.lib, rx.h
    frames = (C.c_uint8 * (32 * 32))()
    res = SyncResult()
    x = np.zeros(1000, np.complex128)
    assert L.b200rx_pass_put(h, x.ctypes.data, 1000) == E_ARG                        # no pass open
    assert L.b200rx_pass_scan(h, 0.0, frames, 32, C.byref(res)) == E_ARG
    assert L.b200rx_pass_open(h) == 0
    t = C.c_uint64()
    sel = (C.c_uint8 * 32)()
    st = (C.c_uint8 * 32)()
    assert L.b200rx_pass_decode(h, sel, None, 600, st, C.byref(t)) == E_ARG           # not scanned yet
    assert L.b200rx_pass_put(h, x.ctypes.data, 1000) == 0
    assert L.b200rx_pass_scan_tagged(h, frames, 32, C.byref(res)) == E_ARG            # fc64 samples are not tagged structs
    assert L.b200rx_pass_scan(h, 0.0, frames, 32, C.byref(res)) == 0 and res.n_frames == 0
    assert L.b200rx_pass_scan(h, 0.0, frames, 32, C.byref(res)) == E_ARG              # scanned already
    assert L.b200rx_pass_put(h, x.ctypes.data, 1000) == E_ARG
    assert L.b200rx_pass_decode(h, sel, None, 600, st, C.byref(t)) == 0 and t.value == 0  # nothing selected: no ticket
    assert L.b200rx_pass_poll(h, 0) == 1 and L.b200rx_pass_wait(h, 0) == 0
    assert L.b200rx_set_sample_format(h, fo.FMT_TAGGED_FC64, 1.0) == 0
    assert L.b200rx_pass_open(h) == 0
    assert L.b200rx_pass_scan(h, 0.0, frames, 32, C.byref(res)) == E_ARG              # tagged streams are not raw captures
    pl = np.zeros((32, 600), np.uint8)
    got = L.b200rx_receive(h, x.ctypes.data, 1000, 0.0, pl.ctypes.data, 600, None, None, st, None, C.byref(res))
    assert got == E_ARG and b"tagged" in L.b200rx_last_error(h)
    rx.close()

    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info()
    for _ in range(6):
        ch = Chain(max_frames=64, max_payload=600, depth=4, max_lag=3)
        ch.run(np.zeros(20000, complex), 4096)
        ch.close()
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < 64 << 20, (free0, free1)   # nothing but allocator slack stays behind
