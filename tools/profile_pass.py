"""Driver for an ncu capture of the small kernels of the two-phase passes (pack_pass_kernel, tag_find_kernel,
tag_frames_kernel): one stream through fun::b200_receiver_chain and through fun::b200_rx.

    ncu --set full --clock-control none -k regex:"pack_pass|tag_find|tag_frames" -c 6 -o out python tools/profile_pass.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from fun_ofdm_b200 import tx  # noqa: E402
from test_gpu_block import Block  # noqa: E402
from test_gpu_chain import Chain  # noqa: E402

rng = np.random.default_rng(1)
n = 256
payloads = rng.integers(0, 256, (n, 1500), dtype=np.uint8)
cap = tx.build_corpus(payloads, np.full(n, 10, np.uint8), snr_db=25.0, lead_in=400, seed=3, threads=os.cpu_count())
x = np.concatenate([np.ascontiguousarray(cap["iq"]), np.zeros(8192, np.complex128)])
ch = Chain(max_frames=512, max_payload=1500)
ch.set_tuning("scan_graph", 0)  # kernels launched one by one so that ncu can name them
got, _, _ = ch.run(x, 1 << 20, max_out=512, stride=1500)
ch.close()
tags = np.zeros(len(x), np.uint8)
tags[cap["frame_off"].astype(np.int64) + 400 + 184] = 4
blk = Block(max_frames=512, max_payload=1500)
got2, _, _ = blk.run(x, tags, 1 << 20, max_out=512, stride=1500)
blk.close()
print("chain payloads", len(got), "block payloads", len(got2))
