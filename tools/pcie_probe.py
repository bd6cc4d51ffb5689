"""Probe: what does the host-buffer entry point leave on the table against a bare pinned H2D copy?
Run on a GPU box:  python tools/pcie_probe.py   (chunk size via B200RX_H2D_CHUNK, one process per value)."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fun_ofdm_b200 import rx as rxmod, tx as txmod  # noqa: E402


def main():
    import fun_ofdm_b200 as fo
    n, plen, rate = 4096, 1500, 10
    payloads = np.random.default_rng(1).integers(0, 256, size=(n, plen), dtype=np.uint8)
    corpus = txmod.build_corpus(payloads, np.full(n, rate, np.uint8), snr_db=25.0, lead_in=0, seed=0xB200)
    iq = np.ascontiguousarray(corpus["iq"]).view(np.float64).reshape(-1)
    iq_samples = iq.size // 2
    lib = fo.load_library()
    r = rxmod.Receiver(0, n, plen)
    sizes = [iq.nbytes, n * 8, n * 4, n * plen, n * 2, n, n]
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
    out = {"chunk": os.environ.get("B200RX_H2D_CHUNK", "512"), "bytes": iq.nbytes}
    host = np.ctypeslib.as_array(C.cast(ptrs[0], C.POINTER(C.c_double)), shape=(iq.size,))
    t = torch.from_numpy(host)
    d = torch.empty(iq.size, dtype=torch.float64, device="cuda")
    for _ in range(3):
        d.copy_(t, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        d.copy_(t, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    out["bare_h2d_ms"] = dt * 1e3
    out["bare_h2d_gbs"] = iq.nbytes / dt / 1e9
    out["torch_sees_pinned"] = bool(t.is_pinned())
    call = lambda: r.decode_batch_ptr(ptrs[0], iq_samples, ptrs[1], ptrs[2], n, ptrs[3], plen, ptrs[4], ptrs[5], ptrs[6])
    for _ in range(3):
        call()
    t0 = time.perf_counter()
    for _ in range(10):
        call()
    dt = (time.perf_counter() - t0) / 10
    out["abi_ms"] = dt * 1e3
    status = np.ctypeslib.as_array(C.cast(ptrs[6], C.POINTER(C.c_uint8)), shape=(n,))
    out["ok"] = int((status == 0).sum())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
