"""Probe: stage durations of one device-resident batch as a function of the number of frames
(how long does the tail chunk of the host-buffer pipeline take?).  B200RX_ACS_LB forces the ACS variant."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fun_ofdm_b200 as fo  # noqa: E402
from fun_ofdm_b200 import tx as txmod  # noqa: E402

plen, rate = 1500, 10
nmax = int(os.environ.get('PROBE_NMAX', '4096'))
payloads = np.random.default_rng(1).integers(0, 256, size=(nmax, plen), dtype=np.uint8)
corpus = txmod.build_corpus(payloads, np.full(nmax, rate, np.uint8), snr_db=25.0, lead_in=0, seed=0xB200)
dev = torch.device("cuda:0")
rx = fo.Receiver(0, nmax, plen)
d_iq = torch.from_numpy(np.ascontiguousarray(corpus["iq"]).view(np.float64)).to(dev)
out = {"lb": os.environ.get("B200RX_ACS_LB", "auto")}
ns = [int(x) for x in os.environ['PROBE_N'].split(',')] if 'PROBE_N' in os.environ else (32, 64, 128, 256, 512, 1024, 2048, 4096)
for n in ns:
    d_l = torch.from_numpy(corpus["lts1"][:n].astype(np.int64)).to(dev)
    d_a = torch.from_numpy(corpus["avail"][:n].astype(np.int32)).to(dev)
    payload = torch.zeros((n, plen), dtype=torch.uint8, device=dev)
    length = torch.zeros(n, dtype=torch.int16, device=dev)
    rt = torch.zeros(n, dtype=torch.uint8, device=dev)
    status = torch.zeros(n, dtype=torch.uint8, device=dev)
    best = None
    for it in range(6):
        rx.decode_batch_dev(d_iq, d_l, d_a, payload, length, rt, status)
        st = rx.stats()
        if it >= 2 and (best is None or st["total_ms"] < best["total_ms"]):
            best = st
    out[n] = [round(best[k], 4) for k in ("frontend_ms", "viterbi_ms", "traceback_ms", "total_ms")] + [int((status == 0).sum())]
print(json.dumps(out))
