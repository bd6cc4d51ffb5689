"""Scratch: host-buffer entry point timing for each sample format / ingest mode (run on a GPU box)."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import fun_ofdm_b200 as fo
from bench import make_corpus

n = 4096
corpus = make_corpus(n, 1500, 10, 25.0, 0xB200, os.cpu_count() or 1)
iq_host = corpus["iq"].view(np.float64)
iq_samples = iq_host.size // 2
rx = fo.Receiver(0, n, 1500)
lib = fo.load_library()


def run(fmt, env, reps=8):
    for k, v in env.items():
        os.environ[k] = str(v)
    if fmt == fo.FMT_FC64:
        wire, scale = iq_host, 1.0
    elif fmt == fo.FMT_FC32:
        wire, scale = iq_host.astype(np.float32), 1.0
    else:
        scale = float(np.max(np.abs(iq_host))) / 30000.0
        wire = np.clip(np.rint(iq_host / scale), -32768, 32767).astype(np.int16)
    sizes = [wire.nbytes, n * 8, n * 4, n * 1500, n * 2, n, n]
    ptrs = []
    for sz in sizes:
        p = C.c_void_p()
        assert lib.b200rx_host_alloc(C.byref(p), sz) == 0
        ptrs.append(p)
    C.memmove(ptrs[0], wire.ctypes.data, wire.nbytes)
    l64 = corpus["lts1"].astype(np.uint64)
    a32 = corpus["avail"].astype(np.uint32)
    C.memmove(ptrs[1], l64.ctypes.data, n * 8)
    C.memmove(ptrs[2], a32.ctypes.data, n * 4)
    rx.set_sample_format(fmt, scale)
    call = lambda: rx.decode_batch_ptr(ptrs[0], iq_samples, ptrs[1], ptrs[2], n, ptrs[3], 1500, ptrs[4], ptrs[5], ptrs[6])
    for _ in range(2):
        call()
    t0 = time.perf_counter()
    for _ in range(reps):
        call()
    ms = (time.perf_counter() - t0) / reps * 1e3
    st = np.ctypeslib.as_array(C.cast(ptrs[6], C.POINTER(C.c_uint8)), shape=(n,))
    ok = int((st == 0).sum())
    for p in ptrs:
        lib.b200rx_host_free(p)
    rx.set_sample_format(fo.FMT_FC64)
    print("fmt %d %-60s %.3f ms  ok %d  %.2f Gbit/s" % (fmt, env, ms, ok, ok * 1500 * 8 / ms / 1e6), flush=True)


for fmt in (0, 1, 2):
    run(fmt, {"B200RX_PULL": 0})
    for ctas in (1, 2, 4):
        run(fmt, {"B200RX_PULL": 1, "B200RX_PULL_CTAS": ctas})
    for ch, chmin in ((512, 64), (2048, 64), (1024, 256)):
        run(fmt, {"B200RX_PULL": 1, "B200RX_PULL_CTAS": 1, "B200RX_H2D_CHUNK": ch, "B200RX_H2D_CHUNK_MIN": chmin})
