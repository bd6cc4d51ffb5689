"""A/B of the add-compare-select kernel variants on BASELINE config 2 (4096 x 1500 B, 54 Mbps, AWGN 25 dB).

For every variant (b200rx_set_tuning): per-stage CUDA-event times of one batch alone and of four batches in one launch
(the saturated regime), and that status / payload bytes are identical to the first variant's.  Prints one JSON line per
variant.  Not a bench value: bench.py is.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fun_ofdm_b200 as fo  # noqa: E402
from fun_ofdm_b200 import tx  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    k = 4
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0xB200)
    payloads = rng.integers(0, 256, (n, 1500), dtype=np.uint8)
    corpus = tx.build_corpus(payloads, np.full(n, 10, np.uint8), snr_db=25.0, lead_in=0, seed=0xB200, threads=os.cpu_count())
    d_iq = torch.from_numpy(corpus["iq"].view(np.float64)).to(dev)
    d_l = torch.from_numpy(corpus["lts1"].astype(np.int64)).to(dev)
    d_a = torch.from_numpy(corpus["avail"].astype(np.int32)).to(dev)
    d_l4, d_a4 = torch.cat([d_l] * k), torch.cat([d_a] * k)

    def outs(m):
        return (torch.zeros((m, 1500), dtype=torch.uint8, device=dev), torch.zeros(m, dtype=torch.int16, device=dev),
                torch.zeros(m, dtype=torch.uint8, device=dev), torch.zeros(m, dtype=torch.uint8, device=dev))

    rx1 = fo.Receiver(0, n, 1500)
    rx4 = fo.Receiver(0, k * n, 1500)
    o1, o4 = outs(n), outs(k * n)
    variants = [dict(acs_gen=2, acs_lb=3, acs_warps=2, acs_rn=1)]
    for lb in (2, 3, 1):
        for rn in (1, 0):
            for w in (1, 4):
                if rn == 0 and w == 4:
                    continue
                variants.append(dict(acs_gen=3, acs_lb=lb, acs_warps=w, acs_rn=rn))
    want = None
    for v in variants:
        line = dict(v)
        for name, rx, l, a, o, reps in (("alone", rx1, d_l, d_a, o1, 5), ("x4", rx4, d_l4, d_a4, o4, 3)):
            for key, val in v.items():
                rx.set_tuning(key, val)
            for _ in range(2):
                rx.decode_batch_dev(d_iq, l, a, *o)
            rx.profile_begin(reps)
            for _ in range(reps):
                rx.decode_batch_dev(d_iq, l, a, *o)
            c, fe, acs, tb = rx.profile_read()
            line[name] = {"fe_ms": round(fe / c, 4), "acs_ms": round(acs / c, 4), "tb_ms": round(tb / c, 4)}
            if name == "alone":
                got = (o[3].cpu().numpy().copy(), o[0].cpu().numpy().copy())
                if want is None:
                    want = got
                    line["frames_ok"] = int((got[0] == 0).sum())
                line["identical"] = bool(np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]))
            else:
                line["acs_ms_per_batch_saturated"] = round(acs / c / k, 4)
                st4 = o[3].cpu().numpy().reshape(k, n)
                line["identical"] = line["identical"] and bool(all(np.array_equal(st4[i], want[0]) for i in range(k)))
        print(json.dumps(line), flush=True)
    rx1.close()
    rx4.close()


if __name__ == "__main__":
    main()
