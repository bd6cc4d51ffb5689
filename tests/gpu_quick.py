"""Scratch driver for gpurun sessions: times the three stages on a config-2 sized batch."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import fun_ofdm_b200 as fo
from oracle import bind
from ofdm_testutil import make_corpus

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ref = bind.ref()
rng = np.random.default_rng(0)
t0 = time.time()
base = make_corpus(ref, rng, [10] * 64, [1500] * 64, snr_db=25, gap=0)
print("corpus 64 frames %.2fs" % (time.time() - t0))
reps = n // 64
iq = np.tile(base["iq"], reps)
flen = len(base["iq"]) // 64
lts1 = np.concatenate([base["lts1"] + r * len(base["iq"]) for r in range(reps)])
avail = np.tile(base["avail"], reps)
dev = torch.device("cuda:0")
rx = fo.Receiver(0, n, 1500)
d_iq = torch.from_numpy(iq.view(np.float64)).to(dev)
d_l = torch.from_numpy(lts1).to(dev)
d_a = torch.from_numpy(avail).to(dev)
payload = torch.zeros((n, 1500), dtype=torch.uint8, device=dev)
length = torch.zeros(n, dtype=torch.int16, device=dev)
rate = torch.zeros(n, dtype=torch.uint8, device=dev)
status = torch.zeros(n, dtype=torch.uint8, device=dev)
for it in range(5):
    rx.decode_batch_dev(d_iq, d_l, d_a, payload, length, rate, status)
    st = rx.stats()
    print(it, {k: (round(v, 4) if isinstance(v, float) else v) for k, v in st.items()})
ok = int((status == 0).sum())
print("ok frames", ok, "of", n, "Gbit/s", ok * 1500 * 8 / (st["total_ms"] * 1e-3) / 1e9)
want = [bytes(p) for p in base["payloads"]]
pl = payload.cpu().numpy()
bad = sum(1 for f in range(n) if status[f] == 0 and bytes(pl[f]) != want[f % 64])
print("payload mismatches vs transmitted:", bad)
