#!/usr/bin/env python
"""Driver for ncu captures of the kernels around the hot path: generates BASELINE config 2 as one raw
capture in HBM with the device generator (tx_frames_kernel, tx_channel_kernel), runs detection + synchronisation +
decode on it (detect_kernel, scan_events_kernel, lts_sync_kernel, build_frames_kernel, frontend_kernel<ROT>), then one
host-buffer call on a pinned copy (pull_kernel).  Prints device times from CUDA events when run without a profiler.

    ncu --set full --clock-control none --import-source on -k regex:"detect|scan_events|lts_sync|build_frames|tx_|pull" \
        -o gpurun_out/aux python tools/profile_aux.py
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fun_ofdm_b200 as fo  # noqa: E402
from fun_ofdm_b200 import tx  # noqa: E402

n, plen, rate, lead = 4096, 1500, 10, 400
dev = torch.device("cuda:0")
rng = np.random.default_rng(0xB200)
payloads = rng.integers(0, 256, (n, plen), dtype=np.uint8)
rates = np.full(n, rate, np.uint8)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


cap = {}


def gen():
    cap.update(tx.build_corpus_dev(payloads, rates, snr_db=25.0, multipath_taps=0, lead_in=lead, seed=0xB201, device=0,
                                   stream=stream.cuda_stream))


gen_ms = timed(gen, reps=3)
iq = cap["iq"]
n_samples = iq.numel() // 2
rx = fo.Receiver(0, n, plen)
rx.set_stream(stream.cuda_stream)
payload = torch.zeros((n, plen), dtype=torch.uint8, device=dev)
length = torch.zeros(n, dtype=torch.int16, device=dev)
r8 = torch.zeros(n, dtype=torch.uint8, device=dev)
status = torch.zeros(n, dtype=torch.uint8, device=dev)
res = {}


def recv():
    res.update(rx.receive_dev(iq, payload, length, r8, status))


recv_ms = timed(recv)
sync_ms = timed(lambda: rx.sync_dev(iq))
ok = int((status == 0).sum())
# the same frames as a genie-tagged batch through the host entry point on a pinned buffer
lib = fo.load_library()
host = iq.cpu().numpy()
p = C.c_void_p()
assert lib.b200rx_host_alloc(C.byref(p), host.nbytes) == 0
C.memmove(p, host.ctypes.data, host.nbytes)
pinned = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(host.size,))
rx.set_stream(None)
lts1 = cap["lts1"].cpu().numpy()
avail = cap["avail"].cpu().numpy()
out = rx.decode_batch(pinned, lts1, avail)
ok_host = int((out[3] == 0).sum())
lib.b200rx_host_free(p)
print({"frames": n, "samples": n_samples, "generator_ms": gen_ms, "generator_gbps": 16 * n_samples / gen_ms / 1e6,
       "receive_ms": recv_ms, "sync_ms": sync_ms, "frames_found": res["n_frames"], "frames_ok": ok, "host_path_ok": ok_host})
