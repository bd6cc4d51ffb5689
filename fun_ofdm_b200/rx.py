"""ctypes binding of include/b200rx.h.  Mirrors the C ABI one to one; no decode logic lives here."""
import ctypes as C
import os

import numpy as np

ST_OK, ST_HDR_PARITY, ST_HDR_RATE, ST_CRC_FAIL, ST_TRUNCATED, ST_TOO_LONG = range(6)
ST_NO_FRAME = 255
RATE_INVALID = 255
FMT_FC64, FMT_FC32, FMT_SC16 = 0, 1, 2
FMT_TAGGED_FC64 = 3  # fun::tagged_sample structs, 24 bytes each (b200rx_pass_scan_tagged, decode entry points)
MAX_INFLIGHT = 3  # B200RX_MAX_INFLIGHT
_FMT_BYTES = {FMT_FC64: 16, FMT_FC32: 8, FMT_SC16: 4, FMT_TAGGED_FC64: 24}

# fun::Rate -> (rate_field, cbps, dbps, bpsc)   reference src/rates.h:52-196
RATE_PARAMS = {
    0: (0xD, 48, 24, 1), 1: (0xE, 48, 32, 1), 2: (0xF, 48, 36, 1),
    3: (0x5, 96, 48, 2), 4: (0x6, 96, 64, 2), 5: (0x7, 96, 72, 2),
    6: (0x9, 192, 96, 4), 7: (0xA, 192, 128, 4), 8: (0xB, 192, 144, 4),
    9: (0x1, 288, 192, 6), 10: (0x3, 288, 216, 6),
}


def num_symbols(rate, length):
    """Data OFDM symbols of a frame (reference src/ppdu.cpp:38-40)."""
    return -(-(16 + 8 * (length + 4) + 6) // RATE_PARAMS[rate][2])


def window_samples(rate, length):
    """Complex samples from the LTS1 tag to the end of the frame: 2 LTS + SIGNAL + data symbols."""
    return 128 + 80 * (1 + num_symbols(rate, length))


class B200RxError(RuntimeError):
    pass


class Limits(C.Structure):
    _fields_ = [("max_frames", C.c_uint32), ("max_payload_bytes", C.c_uint32), ("reserved", C.c_uint32 * 6)]


class Debug(C.Structure):
    _fields_ = [("equalized", C.c_void_p), ("eq_vectors", C.c_uint32),
                ("decoded", C.c_void_p), ("decoded_stride", C.c_uint32),
                ("header_field", C.c_void_p),
                ("depunct", C.c_void_p), ("depunct_stride", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("frontend_ms", C.c_float), ("viterbi_ms", C.c_float), ("traceback_ms", C.c_float),
                ("total_ms", C.c_float), ("trellis_steps", C.c_uint64), ("frames_ok", C.c_uint32),
                ("frames_failed", C.c_uint32), ("payload_bytes", C.c_uint64),
                ("traceback_rewalks", C.c_uint64)]


class SyncResult(C.Structure):
    _fields_ = [("n_events", C.c_uint32), ("n_frames", C.c_uint32), ("overflow", C.c_uint32), ("reserved", C.c_uint32),
                ("last_phase", C.c_double), ("phase_valid", C.c_uint32), ("pad", C.c_uint32)]

    def as_dict(self):
        return {"n_events": int(self.n_events), "n_frames": int(self.n_frames), "overflow": int(self.overflow),
                "last_phase": float(self.last_phase), "phase_valid": bool(self.phase_valid)}


def lib_path():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libb200rx.so")


_lib = None


def load_library():
    """Load libb200rx.so and declare every prototype of include/b200rx.h.  Raises if it is missing:
    the product has no other implementation of the path."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise B200RxError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)" % path)
    L = C.CDLL(path)
    vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
    L.b200rx_create.restype = C.c_int
    L.b200rx_create.argtypes = [C.c_int, C.POINTER(Limits), C.POINTER(vp)]
    L.b200rx_destroy.restype = C.c_int
    L.b200rx_destroy.argtypes = [vp]
    L.b200rx_last_error.restype = C.c_char_p
    L.b200rx_last_error.argtypes = [vp]
    L.b200rx_version.restype = C.c_char_p
    L.b200rx_version.argtypes = []
    L.b200rx_set_stream.restype = C.c_int
    L.b200rx_set_stream.argtypes = [vp, vp]
    L.b200rx_synchronize.restype = C.c_int
    L.b200rx_synchronize.argtypes = [vp]
    L.b200rx_set_sample_format.restype = C.c_int
    L.b200rx_set_sample_format.argtypes = [vp, C.c_int, C.c_double]
    L.b200rx_set_tuning.restype = C.c_int
    L.b200rx_set_tuning.argtypes = [vp, C.c_char_p, C.c_int64]
    L.b200rx_set_pipeline_depth.restype = C.c_int
    L.b200rx_set_pipeline_depth.argtypes = [vp, u32]
    L.b200rx_join.restype = C.c_int
    L.b200rx_join.argtypes = [vp, u32]
    L.b200rx_join_on.restype = C.c_int
    L.b200rx_join_on.argtypes = [vp, u32, vp]
    L.b200rx_host_alloc.restype = C.c_int
    L.b200rx_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    L.b200rx_host_free.restype = C.c_int
    L.b200rx_host_free.argtypes = [vp]
    L.b200rx_decode_batch.restype = C.c_int
    L.b200rx_decode_batch.argtypes = [vp, vp, u64, vp, vp, u32, vp, u32, vp, vp, vp]
    L.b200rx_submit_batch.restype = C.c_int
    L.b200rx_submit_batch.argtypes = [vp, vp, u64, vp, vp, u32, vp, u32, vp, vp, vp, C.POINTER(u64)]
    L.b200rx_wait.restype = C.c_int
    L.b200rx_wait.argtypes = [vp, u64]
    L.b200rx_decode_batch_dev.restype = C.c_int
    L.b200rx_decode_batch_dev.argtypes = [vp, vp, u64, vp, vp, u32, vp, u32, vp, vp, vp, C.POINTER(Debug)]
    L.b200rx_sync_dev.restype = C.c_int
    L.b200rx_sync_dev.argtypes = [vp, vp, u64, C.c_double, vp, vp, vp, vp, C.POINTER(SyncResult)]
    L.b200rx_set_receive_origins.restype = C.c_int
    L.b200rx_set_receive_origins.argtypes = [vp, vp, u32]
    L.b200rx_receive_dev.restype = C.c_int
    L.b200rx_receive_dev.argtypes = [vp, vp, u64, C.c_double, vp, u32, vp, vp, vp, vp, vp, C.POINTER(SyncResult)]
    L.b200rx_receive.restype = C.c_int
    L.b200rx_receive.argtypes = [vp, vp, u64, C.c_double, vp, u32, vp, vp, vp, vp, C.POINTER(SyncResult)]
    L.b200rx_decode_headers.restype = C.c_int
    L.b200rx_decode_headers.argtypes = [vp, vp, u64, vp, vp, u32, vp, vp, vp]
    L.b200rx_viterbi_batch_dev.restype = C.c_int
    L.b200rx_viterbi_batch_dev.argtypes = [vp, vp, u64, vp, u32, u32, vp, u32]
    L.b200rx_get_stats.restype = C.c_int
    L.b200rx_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.b200rx_profile_begin.restype = C.c_int
    L.b200rx_profile_begin.argtypes = [vp, u32]
    L.b200rx_profile_read.restype = C.c_int
    L.b200rx_profile_read.argtypes = [vp, C.POINTER(u32), C.POINTER(C.c_float), C.POINTER(C.c_float),
                                      C.POINTER(C.c_float)]
    L.b200rx_device_counters.restype = C.c_int
    L.b200rx_device_counters.argtypes = [vp, C.POINTER(vp)]
    L.b200rx_copy_counters.restype = C.c_int
    L.b200rx_copy_counters.argtypes = [vp, vp]
    L.b200rx_launch_count.restype = u64
    L.b200rx_launch_count.argtypes = [vp]
    L.b200rx_max_steps.restype = u32
    L.b200rx_max_steps.argtypes = [vp]
    L.b200rx_sample_bytes.restype = C.c_size_t
    L.b200rx_sample_bytes.argtypes = [vp]
    L.b200rx_host_is_pinned.restype = C.c_int
    L.b200rx_host_is_pinned.argtypes = [vp]
    # two-phase passes (streaming callers)
    L.b200rx_pass_open.restype = C.c_int
    L.b200rx_pass_open.argtypes = [vp]
    L.b200rx_pass_put.restype = C.c_int
    L.b200rx_pass_put.argtypes = [vp, vp, u64]
    L.b200rx_pass_scan.restype = C.c_int
    L.b200rx_pass_scan.argtypes = [vp, C.c_double, vp, u32, C.POINTER(SyncResult)]
    L.b200rx_pass_decode.restype = C.c_int
    L.b200rx_pass_decode.argtypes = [vp, vp, vp, u32, vp, C.POINTER(u64)]
    L.b200rx_pass_poll.restype = C.c_int
    L.b200rx_pass_poll.argtypes = [vp, u64]
    L.b200rx_pass_wait.restype = C.c_int
    L.b200rx_pass_wait.argtypes = [vp, u64]
    # several GPUs from one process
    L.b200rx_group_create.restype = C.c_int
    L.b200rx_group_create.argtypes = [vp, u32, C.POINTER(Limits), C.POINTER(vp)]
    L.b200rx_group_destroy.restype = C.c_int
    L.b200rx_group_destroy.argtypes = [vp]
    L.b200rx_group_size.restype = u32
    L.b200rx_group_size.argtypes = [vp]
    L.b200rx_group_handle.restype = vp
    L.b200rx_group_handle.argtypes = [vp, u32]
    L.b200rx_group_last_error.restype = C.c_char_p
    L.b200rx_group_last_error.argtypes = [vp]
    L.b200rx_group_synchronize.restype = C.c_int
    L.b200rx_group_synchronize.argtypes = [vp]
    L.b200rx_group_plan.restype = C.c_int
    L.b200rx_group_plan.argtypes = [vp, vp, u32, vp]
    L.b200rx_group_decode_batch.restype = C.c_int
    L.b200rx_group_decode_batch.argtypes = [vp, vp, u64, vp, vp, u32, vp, u32, vp, vp, vp]
    L.b200rx_group_decode_batch_dev.restype = C.c_int
    L.b200rx_group_decode_batch_dev.argtypes = [vp, vp, vp, vp, vp, vp, vp, u32, vp, vp, vp]
    L.b200rx_gather_status.restype = C.c_int
    L.b200rx_gather_status.argtypes = [vp, vp, u32, vp, vp]
    _lib = L
    return L


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(None)


class Receiver:
    """One b200rx handle (one GPU).  Thin: every method is one C-ABI call."""

    def __init__(self, device=0, max_frames=4096, max_payload_bytes=1500):
        self.lib = load_library()
        self.max_frames = int(max_frames)
        self.max_payload_bytes = int(max_payload_bytes)
        self.device = int(device)
        lim = Limits(self.max_frames, self.max_payload_bytes, (C.c_uint32 * 6)())
        h = C.c_void_p()
        rc = self.lib.b200rx_create(self.device, C.byref(lim), C.byref(h))
        if rc != 0:
            raise B200RxError("b200rx_create failed (%d): %s" % (rc, self.lib.b200rx_last_error(None).decode()))
        self.h = h
        self.fmt = FMT_FC64

    def set_sample_format(self, fmt, sc16_scale=1.0):
        """FMT_FC64 (default), FMT_FC32 (float32 re, im) or FMT_SC16 (int16 re, im; sample = int16 * sc16_scale)."""
        self._check(self.lib.b200rx_set_sample_format(self.h, int(fmt), float(sc16_scale)), "b200rx_set_sample_format")
        self.fmt = int(fmt)

    def _host_samples(self, iq):
        """numpy array in the handle's format -> (contiguous array, number of complex samples)"""
        iq = np.ascontiguousarray(iq)
        want = {FMT_FC64: (np.complex128, np.float64), FMT_FC32: (np.complex64, np.float32), FMT_SC16: (np.int16, np.int16)}[self.fmt]
        if iq.dtype not in want:
            raise B200RxError("samples of dtype %s do not match the handle's sample format %d" % (iq.dtype, self.fmt))
        return iq, iq.nbytes // _FMT_BYTES[self.fmt]

    def _dev_samples(self, iq):
        return int(iq.numel() * iq.element_size()) // _FMT_BYTES[self.fmt]

    def close(self):
        if getattr(self, "h", None):
            self.lib.b200rx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise B200RxError("%s failed (%d): %s" % (what, rc, self.lib.b200rx_last_error(self.h).decode()))

    @property
    def max_steps(self):
        return int(self.lib.b200rx_max_steps(self.h))

    @property
    def launch_count(self):
        return int(self.lib.b200rx_launch_count(self.h))

    def set_stream(self, cuda_stream_ptr):
        self._check(self.lib.b200rx_set_stream(self.h, C.c_void_p(cuda_stream_ptr or None)), "b200rx_set_stream")

    def set_tuning(self, key, value):
        """Implementation knob of the handle (include/b200rx.h, b200rx_set_tuning); never changes results."""
        self._check(self.lib.b200rx_set_tuning(self.h, key.encode(), int(value)), "b200rx_set_tuning")

    def set_pipeline_depth(self, depth):
        self._check(self.lib.b200rx_set_pipeline_depth(self.h, int(depth)), "b200rx_set_pipeline_depth")

    def join(self, calls_back=0):
        self._check(self.lib.b200rx_join(self.h, int(calls_back)), "b200rx_join")

    def join_on(self, calls_back, cuda_stream_ptr):
        self._check(self.lib.b200rx_join_on(self.h, int(calls_back), C.c_void_p(cuda_stream_ptr)), "b200rx_join_on")

    def synchronize(self):
        self._check(self.lib.b200rx_synchronize(self.h), "b200rx_synchronize")

    def stats(self):
        st = Stats()
        self._check(self.lib.b200rx_get_stats(self.h, C.byref(st)), "b200rx_get_stats")
        return {k: getattr(st, k) for k, _ in Stats._fields_}

    def copy_counters(self, dst_ptr):
        """Stream-ordered copy of the most recent call's four counters to device address dst_ptr (b200rx_copy_counters)."""
        self._check(self.lib.b200rx_copy_counters(self.h, C.c_void_p(dst_ptr)), "b200rx_copy_counters")

    def device_counters(self):
        """torch int64[4] view of the handle's device counters {ok, failed, payload bytes, trellis steps}."""
        import torch
        p = C.c_void_p()
        self._check(self.lib.b200rx_device_counters(self.h, C.byref(p)), "b200rx_device_counters")

        class _Arr:
            __cuda_array_interface__ = {"shape": (4,), "typestr": "<i8", "data": (p.value, False), "version": 2}

        return torch.as_tensor(_Arr(), device=torch.device("cuda", self.device))

    def profile_begin(self, slots):
        self._check(self.lib.b200rx_profile_begin(self.h, int(slots)), "b200rx_profile_begin")

    def profile_read(self):
        """-> (calls, frontend_ms_sum, viterbi_ms_sum, traceback_ms_sum)"""
        n, a, b, c = C.c_uint32(), C.c_float(), C.c_float(), C.c_float()
        self._check(self.lib.b200rx_profile_read(self.h, C.byref(n), C.byref(a), C.byref(b), C.byref(c)),
                    "b200rx_profile_read")
        return n.value, a.value, b.value, c.value

    # ---- host buffers (numpy) ----
    def decode_batch(self, iq, lts1_index, avail, payload_stride=None):
        """iq: complex128 or float64 (re, im interleaved) numpy array.  Returns
        (payload[n, stride] u8, length[n] u16, rate[n] u8, status[n] u8)."""
        iq, iq_samples = self._host_samples(iq)
        lts1 = np.ascontiguousarray(lts1_index, dtype=np.uint64)
        av = np.ascontiguousarray(avail, dtype=np.uint32)
        n = len(lts1)
        stride = int(payload_stride or self.max_payload_bytes)
        payload = np.zeros((n, stride), dtype=np.uint8)
        length = np.zeros(n, dtype=np.uint16)
        rate = np.zeros(n, dtype=np.uint8)
        status = np.zeros(n, dtype=np.uint8)
        rc = self.lib.b200rx_decode_batch(self.h, iq.ctypes.data, iq_samples, lts1.ctypes.data, av.ctypes.data, n,
                                          payload.ctypes.data, stride, length.ctypes.data, rate.ctypes.data,
                                          status.ctypes.data)
        self._check(rc, "b200rx_decode_batch")
        return payload, length, rate, status

    def decode_batch_ptr(self, iq_ptr, iq_samples, lts1_ptr, avail_ptr, n, payload_ptr, stride, len_ptr, rate_ptr,
                         status_ptr):
        """Host-buffer entry point on raw addresses (pinned buffers owned by the caller)."""
        rc = self.lib.b200rx_decode_batch(self.h, iq_ptr, iq_samples, lts1_ptr, avail_ptr, n, payload_ptr, stride,
                                          len_ptr, rate_ptr, status_ptr)
        self._check(rc, "b200rx_decode_batch")

    def submit_batch_ptr(self, iq_ptr, iq_samples, lts1_ptr, avail_ptr, n, payload_ptr, stride, len_ptr, rate_ptr,
                         status_ptr):
        """Asynchronous host-buffer entry point on raw addresses; returns the ticket for wait()."""
        t = C.c_uint64()
        rc = self.lib.b200rx_submit_batch(self.h, iq_ptr, iq_samples, lts1_ptr, avail_ptr, n, payload_ptr, stride,
                                          len_ptr, rate_ptr, status_ptr, C.byref(t))
        self._check(rc, "b200rx_submit_batch")
        return t.value

    def wait(self, ticket=0):
        self._check(self.lib.b200rx_wait(self.h, int(ticket)), "b200rx_wait")

    # ---- device buffers (torch tensors on this device) ----
    def decode_batch_dev(self, iq, lts1_index, avail, payload, length, rate, status, debug=None):
        """All arguments are torch CUDA tensors: iq float64 [2*samples] (or complex128), lts1_index int64/uint64
        [n], avail int32 [n], payload uint8 [n, stride], length int16 [n], rate uint8 [n], status uint8 [n].
        Asynchronous on the handle's stream."""
        n = int(lts1_index.numel())
        iq_samples = self._dev_samples(iq)
        dbg = None
        if debug:
            dbg = Debug()
            if debug.get("equalized") is not None:
                dbg.equalized = debug["equalized"].data_ptr()
                dbg.eq_vectors = int(debug["equalized"].shape[1])
            if debug.get("decoded") is not None:
                dbg.decoded = debug["decoded"].data_ptr()
                dbg.decoded_stride = int(debug["decoded"].shape[1])
            if debug.get("header_field") is not None:
                dbg.header_field = debug["header_field"].data_ptr()
            if debug.get("depunct") is not None:
                dbg.depunct = debug["depunct"].data_ptr()
                dbg.depunct_stride = int(debug["depunct"].shape[1])
        rc = self.lib.b200rx_decode_batch_dev(
            self.h, _ptr(iq), iq_samples, _ptr(lts1_index), _ptr(avail), n,
            _ptr(payload), int(payload.shape[1]) if payload is not None else 0,
            _ptr(length), _ptr(rate), _ptr(status), C.byref(dbg) if dbg is not None else None)
        self._check(rc, "b200rx_decode_batch_dev")

    # ---- raw sample streams: frame_detector + timing_sync (+ decode) ----
    def sync_dev(self, iq, phase_in=0.0, tags=None, lts1_index=None, avail=None, phase=None):
        """iq: CUDA tensor float64 [2*n] / complex128 [n].  Optional outputs: tags uint8 [n], lts1_index int64
        [max_frames], avail int32 [max_frames], phase float64 [max_frames].  Returns the summary dict."""
        n = self._dev_samples(iq)
        res = SyncResult()
        rc = self.lib.b200rx_sync_dev(self.h, _ptr(iq), n, float(phase_in), _ptr(tags), _ptr(lts1_index), _ptr(avail),
                                      _ptr(phase), C.byref(res))
        self._check(rc, "b200rx_sync_dev")
        return res.as_dict()

    def receive_dev(self, iq, payload, length, rate, status, lts1_index=None, phase_in=0.0, n_frames=None,
                    wait=True):
        """Raw samples (CUDA tensor) -> payloads of the frames found, in stream order.  Output tensors are sized for
        max_frames; slots beyond the frames found get status ST_NO_FRAME.  wait=True: returns the summary dict once
        the outputs are complete.  wait=False: asynchronous (pipelined over the lanes of set_pipeline_depth); the
        number of frames lands in the optional int32 [1] CUDA tensor `n_frames`."""
        n = self._dev_samples(iq)
        res = SyncResult()
        rc = self.lib.b200rx_receive_dev(self.h, _ptr(iq), n, float(phase_in), _ptr(payload),
                                         int(payload.shape[1]) if payload is not None else 0, _ptr(length), _ptr(rate),
                                         _ptr(status), _ptr(lts1_index), _ptr(n_frames),
                                         C.byref(res) if wait else None)
        self._check(rc, "b200rx_receive_dev")
        return res.as_dict() if wait else None

    def set_receive_origins(self, origins):
        """Work()-buffer origins (chunk start - 160, relative to the next capture) of a chunked reference stream."""
        o = np.ascontiguousarray(origins, dtype=np.int64)
        self._check(self.lib.b200rx_set_receive_origins(self.h, o.ctypes.data, len(o)), "b200rx_set_receive_origins")

    def receive(self, samples, phase_in=0.0):
        """Host samples (complex128 array) -> (list of payload bytes of CRC-OK frames in stream order, info dict with
        per-frame status / length / rate / lts1 and the sync summary).  Mirrors receiver_chain::process_samples
        (receiver_chain.cpp:106-126) for one contiguous capture."""
        iq, n = self._host_samples(samples)
        mf, stride = self.max_frames, max(1, self.max_payload_bytes)
        payload = np.zeros((mf, stride), dtype=np.uint8)
        length = np.zeros(mf, dtype=np.uint16)
        rate = np.zeros(mf, dtype=np.uint8)
        status = np.zeros(mf, dtype=np.uint8)
        lts1 = np.zeros(mf, dtype=np.uint64)
        res = SyncResult()
        rc = self.lib.b200rx_receive(self.h, iq.ctypes.data_as(C.c_void_p), n, float(phase_in),
                                     payload.ctypes.data_as(C.c_void_p), stride, length.ctypes.data_as(C.c_void_p),
                                     rate.ctypes.data_as(C.c_void_p), status.ctypes.data_as(C.c_void_p),
                                     lts1.ctypes.data_as(C.c_void_p), C.byref(res))
        self._check(rc, "b200rx_receive")
        nf = int(res.n_frames)
        out = [bytes(payload[f, : length[f]]) for f in range(nf) if status[f] == ST_OK]
        info = dict(res.as_dict(), status=status[:nf].copy(), length=length[:nf].copy(), rate=rate[:nf].copy(),
                    lts1=lts1[:nf].copy())
        return out, info

    def viterbi_batch_dev(self, symbols, data_bits, max_data_bits, out):
        """symbols uint8 [n, stride] depunctured soft symbols; data_bits int32 [n]; out uint8 [n, out_stride]."""
        n = int(data_bits.numel())
        rc = self.lib.b200rx_viterbi_batch_dev(self.h, _ptr(symbols), int(symbols.shape[1]), _ptr(data_bits),
                                               int(max_data_bits), n, _ptr(out), int(out.shape[1]))
        self._check(rc, "b200rx_viterbi_batch_dev")
