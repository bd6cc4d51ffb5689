"""TEST INFRASTRUCTURE — ctypes bindings for the checker libraries under oracle/.

  * ``Ref``    -> oracle/_ref/libfunref.so : the unmodified reference (bmorgan5/fun_ofdm)
                  compiled by ``make -C oracle ref``; see oracle/ref_harness.cpp.
  * ``Port``   -> oracle/liboracle.so      : the plain-C restatement, oracle/ofdm_oracle.c.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference``
legs may import this module.  The product package (fun_ofdm_b200/) never does.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libfunref.so")
PORT_SO = os.path.join(HERE, "liboracle.so")

_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")

# rate enum -> (rate_field, cbps, dbps, bpsc) — reference rates.h:52-196 (hard-codable integers)
RATES = {
    0: (0xD, 48, 24, 1), 1: (0xE, 48, 32, 1), 2: (0xF, 48, 36, 1),
    3: (0x5, 96, 48, 2), 4: (0x6, 96, 64, 2), 5: (0x7, 96, 72, 2),
    6: (0x9, 192, 96, 4), 7: (0xA, 192, 128, 4), 8: (0xB, 192, 144, 4),
    9: (0x1, 288, 192, 6), 10: (0x3, 288, 216, 6),
}
RATE_NAMES = ["1/2 BPSK", "2/3 BPSK", "3/4 BPSK", "1/2 QPSK", "2/3 QPSK", "3/4 QPSK",
              "1/2 QAM16", "2/3 QAM16", "3/4 QAM16", "2/3 QAM64", "3/4 QAM64"]


def num_symbols(rate, length):
    dbps = RATES[rate][2]
    return -(-(16 + 8 * (length + 4) + 6) // dbps)


def frame_samples(rate, length):
    """Samples in a built frame: 320 preamble + 80 per (SIGNAL + data) symbol."""
    return 320 + 80 * (1 + num_symbols(rate, length))


def window_samples(rate, length):
    """Samples from the LTS1 tag to the end of the frame (the hot path's input window)."""
    return 128 + 80 * (1 + num_symbols(rate, length))


class FrameInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "hdr_ok", "hdr_field", "hdr_parity", "rate_valid", "rate", "length", "nsym", "crc_ok",
        "n_vectors", "payload_from_blocks")]


def _as_iq(x):
    """complex128 array -> float64 view (re, im interleaved), contiguous."""
    x = np.ascontiguousarray(x, dtype=np.complex128)
    return x.view(np.float64)


class _FrameDump:
    """Everything one frame produced on the way through the hot path."""

    def __init__(self, info, eq, soft, deint, depunct, decoded, descrambled, payload):
        self.hdr_ok = bool(info.hdr_ok)
        self.hdr_field = info.hdr_field
        self.hdr_parity = info.hdr_parity
        self.rate_valid = bool(info.rate_valid)
        self.rate = info.rate
        self.length = info.length
        self.nsym = info.nsym
        self.crc_ok = bool(info.crc_ok)
        self.n_vectors = info.n_vectors
        self.eq = eq
        self.soft = soft
        self.deint = deint
        self.depunct = depunct
        self.decoded = decoded
        self.descrambled = descrambled
        self.payload = payload


class _Lib:
    """Shared shape of the two checker libraries: same entry points, prefix differs."""

    prefix = None
    path = None

    def __init__(self):
        if not os.path.exists(self.path):
            raise FileNotFoundError(
                "%s not built: run `make -C oracle %s`" % (self.path, "ref" if self.prefix == "ref" else "port"))
        self.lib = C.CDLL(self.path)
        p = self.prefix
        L = self.lib

        def fn(name, res, args):
            try:
                f = getattr(L, p + "_" + name)
            except AttributeError:  # the port restates the receive path only (no TX)
                return None
            f.restype = res
            f.argtypes = args
            return f

        self._init = fn("init", None, [])
        self._build_frame = fn("build_frame", C.c_int, [_u8p, C.c_int, C.c_int, _f64p, C.c_int])
        self._ppdu_encode = fn("ppdu_encode", C.c_int, [_u8p, C.c_int, C.c_int, _f64p, C.c_int])
        self._conv_encode = fn("conv_encode", None, [_u8p, _u8p, C.c_int])
        self._conv_decode = fn("conv_decode", None, [_u8p, _u8p, C.c_int])
        self._puncture = fn("puncture", C.c_int, [_u8p, C.c_int, C.c_int, _u8p])
        self._depuncture = fn("depuncture", C.c_int, [_u8p, C.c_int, C.c_int, _u8p])
        self._interleave = fn("interleave", C.c_int, [_u8p, C.c_int, _u8p])
        self._deinterleave = fn("deinterleave", C.c_int, [_u8p, C.c_int, _u8p])
        self._modulate = fn("modulate", C.c_int, [_u8p, C.c_int, C.c_int, _f64p])
        self._demodulate = fn("demodulate", C.c_int, [_f64p, C.c_int, C.c_int, _u8p])
        self._fft_forward = fn("fft_forward", None, [_f64p])
        self._crc32 = fn("crc32", C.c_uint32, [_u8p, C.c_int])
        self._parity = fn("parity", C.c_int, [C.c_int])
        self._decode_header = fn("decode_header", C.c_int, [_f64p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)])
        self._decode_data = fn("decode_data", C.c_int, [_f64p, C.c_int, C.c_int, C.c_int, _u8p])
        self._decode_frame = fn("decode_frame", C.c_int, [
            _f64p, C.c_int, C.POINTER(FrameInfo), C.c_void_p, C.c_int,
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p])
        self._decode_batch = fn("decode_batch", C.c_double, [
            _f64p, _i64p, _i32p, C.c_int, _u8p, C.c_int, _i32p, _u8p, C.c_int])
        self._init()

    # ---- TX ----
    def build_frame(self, payload, rate):
        payload = np.ascontiguousarray(np.frombuffer(bytes(payload), dtype=np.uint8))
        n = frame_samples(rate, len(payload))
        out = np.zeros(2 * n, dtype=np.float64)
        pl = payload if len(payload) else np.zeros(1, np.uint8)
        got = self._build_frame(pl, len(payload), rate, out, n)
        assert got == n, (got, n)
        return out.view(np.complex128)

    def ppdu_encode(self, payload, rate):
        payload = np.ascontiguousarray(np.frombuffer(bytes(payload), dtype=np.uint8))
        n = 48 * (1 + num_symbols(rate, len(payload)))
        out = np.zeros(2 * n, dtype=np.float64)
        pl = payload if len(payload) else np.zeros(1, np.uint8)
        got = self._ppdu_encode(pl, len(payload), rate, out, n)
        assert got == n, (got, n)
        return out.view(np.complex128)

    # ---- codec stages ----
    def conv_encode(self, data, data_bits):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        assert len(data) * 8 >= data_bits + 6
        out = np.zeros(2 * (data_bits + 6), dtype=np.uint8)
        self._conv_encode(data, out, data_bits)
        return out

    def conv_decode(self, symbols, data_bits):
        symbols = np.ascontiguousarray(symbols, dtype=np.uint8)
        assert len(symbols) >= 2 * (data_bits + 6)
        out = np.zeros((data_bits + 7) // 8 + 1, dtype=np.uint8)
        self._conv_decode(symbols, out, data_bits)
        return out[: (data_bits + 7) // 8]

    def puncture(self, data, rate):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        out = np.zeros(len(data) + 8, dtype=np.uint8)
        n = self._puncture(data, len(data), rate, out)
        return out[:n]

    def depuncture(self, data, rate):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        out = np.zeros(2 * len(data) + 8, dtype=np.uint8)
        n = self._depuncture(data, len(data), rate, out)
        return out[:n]

    def interleave(self, data):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        out = np.zeros(len(data), dtype=np.uint8)
        self._interleave(data, len(data), out)
        return out

    def deinterleave(self, data):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        out = np.zeros(len(data), dtype=np.uint8)
        self._deinterleave(data, len(data), out)
        return out

    def modulate(self, bits, rate):
        bits = np.ascontiguousarray(bits, dtype=np.uint8)
        n = len(bits) // RATES[rate][3]
        out = np.zeros(2 * n, dtype=np.float64)
        got = self._modulate(bits, len(bits), rate, out)
        assert got == n
        return out.view(np.complex128)

    def demodulate(self, samples, rate):
        iq = _as_iq(samples)
        n = len(iq) // 2
        out = np.zeros(n * RATES[rate][3], dtype=np.uint8)
        got = self._demodulate(iq, n, rate, out)
        assert got == len(out)
        return out

    def fft_forward(self, samples64):
        iq = _as_iq(samples64).copy()
        assert len(iq) == 128
        self._fft_forward(iq)
        return iq.view(np.complex128)

    def crc32(self, data):
        data = np.ascontiguousarray(np.frombuffer(bytes(data), dtype=np.uint8))
        return int(self._crc32(data if len(data) else np.zeros(1, np.uint8), len(data)))

    def parity(self, x):
        return int(self._parity(int(x)))

    def decode_header(self, samples48):
        iq = _as_iq(samples48)
        r, l, n = C.c_int(), C.c_int(), C.c_int()
        ok = self._decode_header(iq, C.byref(r), C.byref(l), C.byref(n))
        return (True, r.value, l.value, n.value) if ok else (False, -1, 0, 0)

    def decode_data(self, samples, rate, length):
        iq = _as_iq(samples)
        out = np.zeros(max(length, 1), dtype=np.uint8)
        ok = self._decode_data(iq, len(iq) // 2, rate, length, out)
        return bool(ok), out[:length]

    # ---- whole frame with every intermediate ----
    def decode_frame(self, window, max_len=4095):
        """window: complex128 samples starting at the LTS1-tagged sample."""
        iq = _as_iq(window)
        n = len(iq) // 2
        nvec_cap = max(0, (n - 128) // 80) + 2
        eq = np.zeros((nvec_cap, 48), dtype=np.complex128)
        cap = nvec_cap * 288
        soft = np.zeros(cap, np.uint8)
        deint = np.zeros(cap, np.uint8)
        depunct = np.zeros(2 * cap, np.uint8)
        decoded = np.zeros(cap // 8 + 64, np.uint8)
        descr = np.zeros(cap // 8 + 64, np.uint8)
        payload = np.zeros(max_len + 8, np.uint8)
        info = FrameInfo()
        self._decode_frame(iq, n, C.byref(info), eq.ctypes.data, nvec_cap, soft.ctypes.data, deint.ctypes.data,
                           depunct.ctypes.data, decoded.ctypes.data, descr.ctypes.data, payload.ctypes.data)
        nv = min(info.n_vectors, nvec_cap)
        full = bool(info.hdr_ok) and info.n_vectors >= 1 + info.nsym
        if full:
            _, cbps, dbps, _ = RATES[info.rate]
            ns, nb = info.nsym, info.nsym * dbps // 8
            d = _FrameDump(info, eq[:nv], soft[: ns * cbps], deint[: ns * cbps], depunct[: 2 * ns * dbps],
                           decoded[:nb], descr[:nb], payload[: info.length] if info.crc_ok else None)
        else:
            d = _FrameDump(info, eq[:nv], None, None, None, None, None, None)
        return d

    def decode_batch(self, iq, lts1_off, n_avail, max_len=4095, threads=1):
        """iq: complex128 stream; per frame the window [lts1_off, lts1_off + n_avail).
        Returns (payload[n, max_len], length[n], status[n], seconds)."""
        iqf = _as_iq(iq)
        lts1_off = np.ascontiguousarray(lts1_off, dtype=np.int64)
        n_avail = np.ascontiguousarray(n_avail, dtype=np.int32)
        n = len(lts1_off)
        payload = np.zeros((n, max_len), dtype=np.uint8)
        length = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.uint8)
        secs = self._decode_batch(iqf, lts1_off, n_avail, n, payload.reshape(-1), max_len, length, status, threads)
        return payload, length, status, secs


class Ref(_Lib):
    prefix = "ref"
    path = REF_SO

    def __init__(self):
        super().__init__()
        L = self.lib
        L.ref_sizeof.restype = C.c_int
        L.ref_sizeof.argtypes = [C.c_int]
        L.ref_sync.restype = C.c_long
        L.ref_sync.argtypes = [_f64p, C.c_long, C.c_int, _f64p, _u8p]
        L.ref_hotpath_stream.restype = C.c_int
        L.ref_hotpath_stream.argtypes = [_f64p, _u8p, C.c_long, C.c_int, _u8p, C.c_int, _i32p, C.c_int]
        L.ref_chain_new.restype = C.c_void_p
        L.ref_chain_new.argtypes = []
        L.ref_chain_process.restype = C.c_int
        L.ref_chain_process.argtypes = [C.c_void_p, _f64p, C.c_int, _u8p, C.c_int, _i32p, C.c_int]

    def table(self, which):
        n = {"preamble": 320, "lts_freq": 64, "lts_time_conj": 64}[which]
        out = np.zeros(2 * n, dtype=np.float64)
        self.lib.ref_table.restype = C.c_int
        self.lib.ref_table.argtypes = [C.c_int, _f64p]
        got = self.lib.ref_table({"preamble": 0, "lts_freq": 1, "lts_time_conj": 2}[which], out)
        assert got == n
        return out.view(np.complex128)

    def sizeof(self, which):
        return self.lib.ref_sizeof({"tagged_sample": 0, "tagged_vector64": 1, "tagged_vector48": 2}[which])

    def sync(self, samples, chunk=4096):
        """frame_detector + timing_sync.  Returns (samples_out, tags) — delayed 160 samples."""
        iq = _as_iq(samples)
        n = len(iq) // 2
        out = np.zeros(2 * n, dtype=np.float64)
        tags = np.zeros(n, dtype=np.uint8)
        got = self.lib.ref_sync(iq, n, chunk, out, tags)
        return out.view(np.complex128)[:got], tags[:got]

    def hotpath_stream(self, samples, tags, chunk=4096, max_len=4095, max_frames=4096):
        iq = _as_iq(samples)
        tags = np.ascontiguousarray(tags, dtype=np.uint8)
        payload = np.zeros((max_frames, max_len), dtype=np.uint8)
        length = np.zeros(max_frames, dtype=np.int32)
        n = self.lib.ref_hotpath_stream(iq, tags, len(tags), chunk, payload.reshape(-1), max_len, length, max_frames)
        n = min(n, max_frames)
        return [bytes(payload[i, : length[i]]) for i in range(n)]

    def chain_new(self):
        return self.lib.ref_chain_new()

    def chain_process(self, chain, samples, max_len=4095, max_frames=256):
        iq = _as_iq(samples)
        payload = np.zeros((max_frames, max_len), dtype=np.uint8)
        length = np.zeros(max_frames, dtype=np.int32)
        n = self.lib.ref_chain_process(chain, iq, len(iq) // 2, payload.reshape(-1), max_len, length, max_frames)
        n = min(n, max_frames)
        return [bytes(payload[i, : length[i]]) for i in range(n)]


    def chain_run(self, chain, samples, chunk, drain=8, max_len=4095, max_frames=8192):
        """The whole test_sim feed loop in native code (ref_harness.cpp: ref_chain_run).  Returns (payloads, seconds)."""
        iq = _as_iq(samples)
        payload = np.zeros((max_frames, max_len), dtype=np.uint8)
        length = np.zeros(max_frames, dtype=np.int32)
        sec = np.zeros(1, dtype=np.float64)
        fn = self.lib.ref_chain_run
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_long, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        n = fn(chain, iq.ctypes.data, len(iq) // 2, int(chunk), int(drain), payload.ctypes.data, max_len, length.ctypes.data,
               max_frames, sec.ctypes.data)
        n = min(n, max_frames)
        return [bytes(payload[i, : length[i]]) for i in range(n)], float(sec[0])


class Port(_Lib):
    prefix = "orc"
    path = PORT_SO


_ref = None
_port = None


def ref():
    global _ref
    if _ref is None:
        _ref = Ref()
    return _ref


def port():
    global _port
    if _port is None:
        _port = Port()
    return _port


def have_ref():
    return os.path.exists(REF_SO)
