"""Experiment: share of the chunks of b200rx_submit_batch that go through the DMA engine while the rest is pulled by the
GPU (pull_mode k >= 2: every k-th chunk by DMA), config 2, pinned fc64 host buffers, three calls in flight."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fun_ofdm_b200 as fo  # noqa: E402
from fun_ofdm_b200 import tx  # noqa: E402

n, plen = 4096, 1500
rng = np.random.default_rng(1)
payloads = rng.integers(0, 256, (n, plen), dtype=np.uint8)
corpus = tx.build_corpus(payloads, np.full(n, 10, np.uint8), snr_db=25.0, lead_in=0, seed=0xB200, threads=os.cpu_count())
iq = np.ascontiguousarray(corpus["iq"]).view(np.float64)
lib = fo.load_library()
rx = fo.Receiver(0, n, plen)
sizes = [iq.nbytes, n * 8, n * 4] + [n * plen, n * 2, n, n] * 3
ptrs = []
for sz in sizes:
    p = C.c_void_p()
    assert lib.b200rx_host_alloc(C.byref(p), sz) == 0
    ptrs.append(p)
C.memmove(ptrs[0], iq.ctypes.data, iq.nbytes)
l64 = corpus["lts1"].astype(np.uint64)
a32 = corpus["avail"].astype(np.uint32)
C.memmove(ptrs[1], l64.ctypes.data, n * 8)
C.memmove(ptrs[2], a32.ctypes.data, n * 4)
for chunk in (1024, 512, 256):
    for mode in (0, 1, 2, 3, 4, 6, 8):
        rx.set_tuning("h2d_chunk", chunk)
        rx.set_tuning("h2d_chunk_min", min(256, chunk))
        rx.set_tuning("pull_mode", mode)

        def run(steps):
            tickets = []
            for j in range(steps):
                o = ptrs[3 + 4 * (j % 3): 7 + 4 * (j % 3)]
                tickets.append(rx.submit_batch_ptr(ptrs[0], iq.size // 2, ptrs[1], ptrs[2], n, o[0], plen, o[1], o[2], o[3]))
                if len(tickets) > 3:
                    rx.wait(tickets.pop(0))
            rx.wait(0)
        run(4)
        t0 = time.perf_counter()
        run(12)
        dt = (time.perf_counter() - t0) / 12
        st = np.ctypeslib.as_array(C.cast(ptrs[6], C.POINTER(C.c_uint8)), shape=(n,))
        print("chunk %4d pull_mode %d: %.3f ms per step, %.2f Gbit/s, ok %d" % (chunk, mode, 1e3 * dt, n * plen * 8 / dt / 1e9, int((st == 0).sum())), flush=True)
rx.close()
