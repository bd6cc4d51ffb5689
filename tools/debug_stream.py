"""Scratch: locate a payload-sequence mismatch found by stress_receive.py.  usage: debug_stream.py seed index"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import fun_ofdm_b200 as fo  # noqa: E402
from oracle import bind  # noqa: E402
from stress_receive import make_stream  # noqa: E402
from test_gpu_chain import _reference_chain  # noqa: E402

seed, index = int(sys.argv[1]), int(sys.argv[2])
extra_draws = sys.argv[3] if len(sys.argv) > 3 else "stress"
ref = bind.ref()
rng = np.random.default_rng(seed)
for s in range(index + 1):
    x, snr, nf = make_stream(ref, rng)
    if s < index and extra_draws == "stress2":
        rng.choice([333, 1000, 4096, 4096, 20000])
    elif s < index and extra_draws == "stress":
        pos = 0
        while pos < len(x):  # stress_receive draws chunk sizes from the same generator
            pos += int(rng.choice([333, 1000, 4096, 20000]))
    elif s < index and extra_draws == "pytest":
        pos = 0
        while pos < len(x):
            pos += int(rng.choice([333, 1000, 4096, 20000]))
print("stream", index, "samples", len(x), "snr", snr, "frames sent", nf)
want = _reference_chain(ref, x, 4096)
chain = ref.chain_new()
per_call = []
for c0 in range(0, len(x), 4096):
    per_call.append([len(p) for p in ref.chain_process(chain, x[c0: c0 + 4096])])
for _ in range(8):
    per_call.append([len(p) for p in ref.chain_process(chain, np.zeros(4096, complex))])
print("reference chain, payload lengths returned by each 4096-sample call:", per_call)
rx = fo.Receiver(0, 256, 4095)
xx = np.concatenate([x, np.zeros(400, complex)])
got, info = rx.receive(xx)
print("reference payload lengths", [len(p) for p in want])
print("gpu       payload lengths", [len(p) for p in got])
print("gpu frames: lts1", info["lts1"].tolist())
print("gpu status", info["status"].tolist(), "len", info["length"].tolist(), "rate", info["rate"].tolist())
# tags
out, tags = ref.sync(xx, chunk=4096)
t = np.zeros(len(xx), np.uint8)
t[: len(xx) - 160] = tags[160: len(xx)]
print("ref LTS1 tags at", np.nonzero(t == 4)[0].tolist())
print("ref STS_END  at", np.nonzero(t == 2)[0].tolist())
dev = torch.device("cuda:0")
d = torch.from_numpy(xx.view(np.float64)).to(dev)
gt = torch.zeros(len(xx), dtype=torch.uint8, device=dev)
res = rx.sync_dev(d, 0.0, gt)
g = gt.cpu().numpy()
print("gpu LTS1 tags at", np.nonzero(g == 4)[0].tolist())
print("gpu STS_END  at", np.nonzero(g == 2)[0].tolist())
# the reference's four hot-path blocks on its own tagged stream
hp = ref.hotpath_stream(out, tags, chunk=4096)
print("ref hotpath_stream lengths", [len(p) for p in hp])
for chunk in (333, 1000, 20000, 1 << 20):
    w2 = _reference_chain(ref, x, chunk)
    print("reference chain with chunk", chunk, [len(p) for p in w2], "same as 4096:", w2 == want)
sys.stdout.flush()
os._exit(0)
