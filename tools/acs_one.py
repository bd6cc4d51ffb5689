"""One ACS variant on config 2 for ncu: python tools/acs_one.py gen lb warps rn [frames] [reps]."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fun_ofdm_b200 as fo  # noqa: E402
from fun_ofdm_b200 import tx  # noqa: E402

gen, lb, warps, rn = (int(x) for x in sys.argv[1:5])
n = int(sys.argv[5]) if len(sys.argv) > 5 else 4096
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 2
base = min(n, 4096)
dev = torch.device("cuda:0")
rng = np.random.default_rng(0xB200)
payloads = rng.integers(0, 256, (base, 1500), dtype=np.uint8)
corpus = tx.build_corpus(payloads, np.full(base, 10, np.uint8), snr_db=25.0, lead_in=0, seed=0xB200, threads=os.cpu_count())
k = n // base
d_iq = torch.from_numpy(corpus["iq"].view(np.float64)).to(dev)
d_l = torch.cat([torch.from_numpy(corpus["lts1"].astype(np.int64)).to(dev)] * k)
d_a = torch.cat([torch.from_numpy(corpus["avail"].astype(np.int32)).to(dev)] * k)
rx = fo.Receiver(0, n, 1500)
for key, val in (("acs_gen", gen), ("acs_lb", lb), ("acs_warps", warps), ("acs_rn", rn)):
    rx.set_tuning(key, val)
o = (torch.zeros((n, 1500), dtype=torch.uint8, device=dev), torch.zeros(n, dtype=torch.int16, device=dev),
     torch.zeros(n, dtype=torch.uint8, device=dev), torch.zeros(n, dtype=torch.uint8, device=dev))
for _ in range(reps):
    rx.decode_batch_dev(d_iq, d_l, d_a, *o)
rx.synchronize()
print("ok frames", int((o[3] == 0).sum()), rx.stats())
rx.close()
