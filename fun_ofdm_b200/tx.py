"""ctypes binding of include/b200tx.h (host-side synthetic corpus generator)."""
import ctypes as C
import os

import numpy as np

from .rx import B200RxError, num_symbols

LTS1_OFFSET = 184  # where timing_sync puts the LTS1 tag relative to the frame start (timing_sync.cpp:105)


class Channel(C.Structure):
    _fields_ = [("snr_db", C.c_double), ("multipath_taps", C.c_uint32), ("lead_in", C.c_uint32),
                ("seed", C.c_uint64), ("n_threads", C.c_uint32), ("reserved", C.c_uint32)]


def host_lib_path():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libb200host.so")


_lib = None


def load_host_library():
    global _lib
    if _lib is not None:
        return _lib
    path = host_lib_path()
    if not os.path.exists(path):
        raise B200RxError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`" % path)
    L = C.CDLL(path)
    vp = C.c_void_p
    L.b200tx_frame_samples.restype = C.c_int
    L.b200tx_frame_samples.argtypes = [C.c_int, C.c_int]
    L.b200tx_num_symbols.restype = C.c_int
    L.b200tx_num_symbols.argtypes = [C.c_int, C.c_int]
    L.b200tx_build_frame.restype = C.c_int
    L.b200tx_build_frame.argtypes = [vp, C.c_int, C.c_int, vp]
    L.b200tx_ppdu_encode.restype = C.c_int
    L.b200tx_ppdu_encode.argtypes = [vp, C.c_int, C.c_int, vp]
    L.b200tx_preamble.restype = None
    L.b200tx_preamble.argtypes = [vp]
    L.b200tx_build_batch.restype = C.c_int
    L.b200tx_build_batch.argtypes = [vp, vp, vp, vp, C.c_uint32, vp, vp, C.POINTER(Channel)]
    _lib = L
    return L


def frame_samples(rate, length):
    return 320 + 80 * (1 + num_symbols(rate, length))


def build_frame(payload, rate):
    """frame_builder::build_frame equivalent -> complex128 array."""
    L = load_host_library()
    pl = np.frombuffer(bytes(payload), dtype=np.uint8)
    n = frame_samples(rate, len(pl))
    out = np.zeros(n, dtype=np.complex128)
    buf = np.ascontiguousarray(pl) if len(pl) else np.zeros(1, np.uint8)
    got = L.b200tx_build_frame(buf.ctypes.data, len(pl), rate, out.ctypes.data)
    if got != n:
        raise B200RxError("b200tx_build_frame returned %d" % got)
    return out


def ppdu_encode(payload, rate):
    L = load_host_library()
    pl = np.frombuffer(bytes(payload), dtype=np.uint8)
    n = 48 * (1 + num_symbols(rate, len(pl)))
    out = np.zeros(n, dtype=np.complex128)
    buf = np.ascontiguousarray(pl) if len(pl) else np.zeros(1, np.uint8)
    got = L.b200tx_ppdu_encode(buf.ctypes.data, len(pl), rate, out.ctypes.data)
    if got != n:
        raise B200RxError("b200tx_ppdu_encode returned %d" % got)
    return out


def preamble():
    out = np.zeros(320, dtype=np.complex128)
    load_host_library().b200tx_preamble(out.ctypes.data)
    return out


def build_corpus(payloads, rates, snr_db=25.0, multipath_taps=0, lead_in=0, seed=0xB200, threads=None, out=None):
    """Frames + channel into one stream.

    payloads: list of bytes (or a uint8 [n, len] array); rates: per-frame fun::Rate values.
    Returns dict(iq complex128 [total], lts1 uint64 [n], avail uint32 [n], frame_off uint64 [n]).
    Frame f occupies [frame_off[f] + lead_in, + frame_samples); its LTS1 tag is 184 samples in.
    """
    L = load_host_library()
    n = len(payloads)
    rates = np.ascontiguousarray(rates, dtype=np.uint8)
    if isinstance(payloads, np.ndarray) and payloads.ndim == 2:
        lengths = np.full(n, payloads.shape[1], dtype=np.uint32)
        blob = np.ascontiguousarray(payloads, dtype=np.uint8).reshape(-1)
        poff = (np.arange(n, dtype=np.uint64) * np.uint64(payloads.shape[1])).astype(np.uint64)
    else:
        lengths = np.array([len(p) for p in payloads], dtype=np.uint32)
        poff = np.zeros(n, dtype=np.uint64)
        poff[1:] = np.cumsum(lengths[:-1], dtype=np.uint64)
        blob = np.frombuffer(b"".join(bytes(p) for p in payloads) + b"\0", dtype=np.uint8)
    ns = np.array([frame_samples(int(r), int(l)) for r, l in zip(rates, lengths)], dtype=np.uint64)
    span = ns + np.uint64(lead_in)
    off = np.zeros(n, dtype=np.uint64)
    off[1:] = np.cumsum(span[:-1], dtype=np.uint64)
    total = int(span.sum())
    if out is None:
        out = np.empty(total, dtype=np.complex128)
    assert out.dtype == np.complex128 and out.size >= total
    ch = Channel(float(snr_db if snr_db is not None else 1000.0), int(multipath_taps), int(lead_in), int(seed),
                 int(threads or os.cpu_count() or 1), 0)
    rc = L.b200tx_build_batch(blob.ctypes.data, poff.ctypes.data, lengths.ctypes.data, rates.ctypes.data, n,
                              out.ctypes.data, off.ctypes.data, C.byref(ch))
    if rc != 0:
        raise B200RxError("b200tx_build_batch failed (%d)" % rc)
    lts1 = off + np.uint64(lead_in + LTS1_OFFSET)
    avail = (ns - np.uint64(LTS1_OFFSET)).astype(np.uint32)
    return dict(iq=out[:total], lts1=lts1, avail=avail, frame_off=off, lengths=lengths, rates=rates)


def build_corpus_dev(payloads, rates, snr_db=25.0, multipath_taps=0, lead_in=0, seed=0xB200, device=0, stream=None):
    """build_corpus on the GPU (b200tx_build_batch_dev in libb200rx.so, fun_ofdm_b200/csrc/txgen.cu): the stream is
    written straight into HBM.  Same arguments, same layout, same seeds as build_corpus; returns torch CUDA tensors:
    dict(iq float64 [2 * total], lts1 int64 [n], avail int32 [n]) plus the numpy frame_off / lengths / rates."""
    import torch

    from . import rx as _rx
    L = _rx.load_library()
    fn = L.b200tx_build_batch_dev
    fn.restype = C.c_int
    fn.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                   C.POINTER(Channel)]
    n = len(payloads)
    rates = np.ascontiguousarray(rates, dtype=np.uint8)
    if isinstance(payloads, np.ndarray) and payloads.ndim == 2:
        lengths = np.full(n, payloads.shape[1], dtype=np.uint32)
        blob = np.ascontiguousarray(payloads, dtype=np.uint8).reshape(-1)
        poff = (np.arange(n, dtype=np.uint64) * np.uint64(payloads.shape[1])).astype(np.uint64)
    else:
        lengths = np.array([len(p) for p in payloads], dtype=np.uint32)
        poff = np.zeros(n, dtype=np.uint64)
        poff[1:] = np.cumsum(lengths[:-1], dtype=np.uint64)
        blob = np.frombuffer(b"".join(bytes(p) for p in payloads) + b"\0", dtype=np.uint8)
    if blob.size == 0:
        blob = np.zeros(1, np.uint8)
    ns = np.array([frame_samples(int(r), int(l)) for r, l in zip(rates, lengths)], dtype=np.uint64)
    span = ns + np.uint64(lead_in)
    off = np.zeros(n, dtype=np.uint64)
    off[1:] = np.cumsum(span[:-1], dtype=np.uint64)
    total = int(span.sum())
    dev = torch.device("cuda", device)
    d_blob = torch.from_numpy(blob.copy()).to(dev)
    d_poff = torch.from_numpy(poff.astype(np.int64)).to(dev)
    d_len = torch.from_numpy(lengths.astype(np.int32)).to(dev)
    d_rates = torch.from_numpy(rates.copy()).to(dev)
    d_off = torch.from_numpy(off.astype(np.int64)).to(dev)
    iq = torch.empty(2 * total, dtype=torch.float64, device=dev)
    ch = Channel(float(snr_db if snr_db is not None else 1000.0), int(multipath_taps), int(lead_in), int(seed), 0, 0)
    s = stream if stream is not None else torch.cuda.current_stream(dev).cuda_stream
    rc = fn(device, C.c_void_p(s), d_blob.data_ptr(), d_poff.data_ptr(), d_len.data_ptr(), d_rates.data_ptr(), n,
            iq.data_ptr(), d_off.data_ptr(), C.byref(ch))
    if rc != 0:
        raise B200RxError("b200tx_build_batch_dev failed (%d)" % rc)
    torch.cuda.current_stream(dev).synchronize() if stream is None else None
    lts1 = torch.from_numpy((off + np.uint64(lead_in + LTS1_OFFSET)).astype(np.int64)).to(dev)
    avail = torch.from_numpy((ns - np.uint64(LTS1_OFFSET)).astype(np.int32)).to(dev)
    return dict(iq=iq, lts1=lts1, avail=avail, frame_off=off, lengths=lengths, rates=rates, keep=(d_blob, d_poff, d_len, d_rates, d_off))
