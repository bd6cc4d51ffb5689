"""GPU: fun::b200_rx (the C++ block that replaces fft_symbols..frame_decoder) against the reference's own
four blocks on the same tagged_sample stream, chunk by chunk like receiver_chain::process_samples."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Block:
    def __init__(self, max_frames=64, max_payload=4095, lib_path=None, depth=None, max_lag=3):
        self.lib = C.CDLL(lib_path or os.path.join(ROOT, "fun_ofdm_b200", "lib", "libb200host.so"))
        self.lib.b200host_rx_block_new.restype = C.c_void_p
        self.lib.b200host_rx_block_new.argtypes = [C.c_int, C.c_uint, C.c_uint]
        self.lib.b200host_rx_block_work.restype = C.c_int
        self.lib.b200host_rx_block_work.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p,
                                                    C.c_int, C.c_void_p, C.c_int]
        self.lib.b200host_rx_block_delete.argtypes = [C.c_void_p]
        self.lib.b200host_rx_block_counters.argtypes = [C.c_void_p, C.c_void_p]
        self.lib.b200host_rx_block_new2.restype = C.c_void_p
        self.lib.b200host_rx_block_new2.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_uint, C.c_uint]
        self.lib.b200host_rx_block_run.restype = C.c_int
        self.lib.b200host_rx_block_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_long, C.c_void_p, C.c_int,
                                                   C.c_void_p, C.c_int, C.c_void_p]
        if depth is None:
            self.h = self.lib.b200host_rx_block_new(0, max_frames, max_payload)
        else:
            self.h = self.lib.b200host_rx_block_new2(0, max_frames, max_payload, depth, max_lag)
        assert self.h, "b200_rx block could not be created (no GPU?)"

    def run(self, samples, tags, chunk, max_out=4096, stride=4095):
        """All rounds + flush in native code (host_capi.cpp: b200host_rx_block_run): (payloads, seconds of the rounds,
        seconds including the flush); the per-round std::vector<tagged_sample> are built before the clock starts."""
        iq = np.ascontiguousarray(samples, dtype=np.complex128).view(np.float64)
        tags = np.ascontiguousarray(tags, dtype=np.uint8)
        payload = np.zeros((max_out, stride), np.uint8)
        length = np.zeros(max_out, np.int32)
        sec = np.zeros(2, np.float64)
        n = self.lib.b200host_rx_block_run(self.h, iq.ctypes.data, tags.ctypes.data, len(tags), chunk, payload.ctypes.data, stride,
                                           length.ctypes.data, max_out, sec.ctypes.data)
        assert n <= max_out
        return [bytes(payload[i, : length[i]]) for i in range(n)], float(sec[0]), float(sec[1])

    def work(self, samples, tags, flush=False, max_out=512, stride=4095):
        iq = np.ascontiguousarray(samples, dtype=np.complex128).view(np.float64)
        tags = np.ascontiguousarray(tags, dtype=np.uint8)
        payload = np.zeros((max_out, stride), np.uint8)
        length = np.zeros(max_out, np.int32)
        n = self.lib.b200host_rx_block_work(self.h, iq.ctypes.data, tags.ctypes.data, len(tags), int(flush),
                                            payload.ctypes.data, stride, length.ctypes.data, max_out)
        return [bytes(payload[i, : length[i]]) for i in range(min(n, max_out))]

    def counters(self):
        c = np.zeros(5, np.uint64)
        self.lib.b200host_rx_block_counters(self.h, c.ctypes.data)
        return dict(zip(["seen", "headers_bad", "ok", "crc_fail", "abandoned"], (int(x) for x in c)))

    def close(self):
        self.lib.b200host_rx_block_delete(self.h)


def _stream(ref, rng, rates, lengths, snr_db, gap=600):
    chunks, payloads = [np.zeros(400, complex)], []
    for rate, length in zip(rates, lengths):
        pl = rng.integers(0, 256, length, dtype=np.uint8).tobytes()
        f = ref.build_frame(pl, rate)
        payloads.append(pl)
        chunks += [f, np.zeros(gap, complex)]
    chunks.append(np.zeros(8192, complex))
    x = np.concatenate(chunks)
    if snr_db is not None:
        p = 0.0127  # mean |x|^2 of a frame body
        sig = np.sqrt(p / 10 ** (snr_db / 10.0) / 2.0)
        x = x + sig * (rng.standard_normal(len(x)) + 1j * rng.standard_normal(len(x)))
    return x, payloads


@pytest.mark.parametrize("snr", [None, 25])
def test_block_matches_reference_blocks_on_synced_stream(ref, snr):
    rng = np.random.default_rng(17 if snr is None else 18)
    rates = [10, 8, 0, 5, 10, 3, 9, 6, 10, 2, 10, 10]
    lengths = [1500, 300, 40, 700, 64, 1000, 1499, 255, 0, 120, 1500, 333]
    x, payloads = _stream(ref, rng, rates, lengths, snr)
    samples, tags = ref.sync(x, chunk=4096)          # reference frame_detector + timing_sync (the boundary producer)
    assert (tags == 4).sum() >= len(rates) - 1        # LTS1 tags
    want = ref.hotpath_stream(samples, tags, chunk=4096)
    blk = Block()
    got = []
    for s in range(0, len(tags), 4096):
        got += blk.work(samples[s: s + 4096], tags[s: s + 4096])
    got += blk.work(np.zeros(1, complex), np.zeros(1, np.uint8), flush=True)
    c = blk.counters()
    blk.close()
    assert got == want, (len(got), len(want), c)
    assert len(got) >= len(rates) - 2
    if snr is None:
        assert got == [p for p in payloads if p in got]


def test_block_genie_tags_and_frame_cut_short(ref):
    """Genie tags (LTS1 at frame start + 184); a frame interrupted by the next LTS1 is abandoned, like the
    reference (whose decoder restarts on the new valid SIGNAL, frame_decoder.cpp:86)."""
    rng = np.random.default_rng(3)
    f1 = ref.build_frame(rng.integers(0, 256, 800, dtype=np.uint8).tobytes(), 10)
    pl2 = rng.integers(0, 256, 90, dtype=np.uint8).tobytes()
    f2 = ref.build_frame(pl2, 6)
    x = np.concatenate([f1[: len(f1) // 2], f2, np.zeros(500, complex)])
    tags = np.zeros(len(x), np.uint8)
    tags[184] = 4
    tags[184 + 64] = 5
    tags[len(f1) // 2 + 184] = 4
    tags[len(f1) // 2 + 184 + 64] = 5
    want = ref.hotpath_stream(x, tags, chunk=1000)
    blk = Block()
    got = []
    for s in range(0, len(x), 1000):
        got += blk.work(x[s: s + 1000], tags[s: s + 1000])
    got += blk.work(np.zeros(1, complex), np.zeros(1, np.uint8), flush=True)  # payloads surface up to max_lag rounds late
    c = blk.counters()
    blk.close()
    assert got == want == [pl2]
    assert c["abandoned"] == 1 and c["ok"] == 1


@pytest.mark.parametrize("depth,lag", [(1, 0), (4, 3), (8, 7)])
def test_block_pipelined_rounds_native_loop(ref, depth, lag):
    """The block inside the reference's round structure (input_buffer swapped in, work(), output_buffer read), native loop,
    with 1, 4 and 8 passes in flight: the tagged_sample structs go to the GPU as they are (24 bytes each, tags found on
    the device); payload sequence = the reference's four blocks on the same tagged stream."""
    rng = np.random.default_rng(808 + depth)
    rates = [int(r) for r in rng.integers(0, 11, 40)]
    lengths = [int(v) for v in rng.integers(0, 1000, 40)]
    x, payloads = _stream(ref, rng, rates, lengths, 24, gap=500)
    for chunk in (4096, 30000):
        samples, tags = ref.sync(x, chunk=chunk)
        n_whole = (len(tags) // chunk) * chunk  # whole rounds only (the reference's double delivery)
        samples, tags = samples[:n_whole], tags[:n_whole]
        want = ref.hotpath_stream(samples, tags, chunk=chunk)
        blk = Block(max_frames=128, depth=depth, max_lag=lag)
        got, t_rounds, t_all = blk.run(samples, tags, chunk)
        c = blk.counters()
        blk.close()
        assert got == want, (depth, lag, chunk, len(got), len(want), c)
        assert len(got) >= 25 and c["ok"] == len(got), c
