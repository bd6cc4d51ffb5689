#!/usr/bin/env python
"""Randomised parity stress of the hot path: batches of frames with random rates (all 11), lengths (0..4095), SNR (2..35 dB),
multipath (0..8 taps) and randomly shortened windows go through b200rx_decode_batch_dev and through the reference's own
four blocks (oracle/_ref, one private instance per host thread).  For every frame: delivered / not delivered must agree and
delivered payloads must be byte-identical; a sample of the frames is compared in full (status code, LENGTH, rate,
descrambled bytes of CRC failures, depunctured soft symbols) against the single-frame checker.

    python tools/stress_decode.py [n_batches] [frames_per_batch] [seed]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fun_ofdm_b200 as fo  # noqa: E402
from fun_ofdm_b200 import tx  # noqa: E402
from oracle import bind  # noqa: E402


def main():
    n_batches = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    ref = bind.ref()
    rng = np.random.default_rng(seed)
    dev = torch.device("cuda:0")
    rx = fo.Receiver(0, n, 4095)
    threads = os.cpu_count() or 1
    tot = dict(frames=0, delivered=0, verdict_mismatch=0, payload_mismatch=0, full_checked=0, full_mismatch=0)
    for b in range(n_batches):
        snr = float(rng.uniform(2, 35))
        taps = int(rng.choice([0, 0, 2, 4, 8]))
        rates = rng.integers(0, 11, n).astype(np.uint8)
        kind = rng.random(n)
        lengths = np.where(kind < 0.5, rng.integers(0, 200, n), np.where(kind < 0.9, rng.integers(0, 1600, n), rng.integers(0, 4096, n)))
        payloads = [rng.integers(0, 256, int(m), dtype=np.uint8).tobytes() for m in lengths]
        c = tx.build_corpus(payloads, rates, snr_db=snr, multipath_taps=taps, lead_in=0, seed=int(rng.integers(1 << 30)), threads=threads)
        avail = c["avail"].astype(np.int32).copy()
        cut = np.nonzero(rng.random(n) < 0.03)[0]
        for f in cut:
            avail[f] = int(rng.integers(0, avail[f] + 1))
        lts1 = c["lts1"].astype(np.int64)
        d_iq = torch.from_numpy(c["iq"].view(np.float64)).to(dev)
        d_l = torch.from_numpy(lts1).to(dev)
        d_a = torch.from_numpy(avail).to(dev)
        payload = torch.zeros((n, 4095), dtype=torch.uint8, device=dev)
        length = torch.zeros(n, dtype=torch.int16, device=dev)
        rate = torch.zeros(n, dtype=torch.uint8, device=dev)
        status = torch.full((n,), 99, dtype=torch.uint8, device=dev)
        dbg = dict(depunct=torch.zeros((n, 2 * rx.max_steps), dtype=torch.uint8, device=dev))
        rx.decode_batch_dev(d_iq, d_l, d_a, payload, length, rate, status, dbg)
        rx.synchronize()
        st = status.cpu().numpy()
        ln = length.cpu().numpy().astype(np.uint16).astype(np.int64)
        pl = payload.cpu().numpy()
        rt = rate.cpu().numpy()
        w_pl, w_len, w_st, _ = ref.decode_batch(c["iq"], lts1, avail, max_len=4095, threads=threads)
        ok_g, ok_r = st == 0, w_st == 0
        # a shortened window can leave the reference's frame_decoder without input, and its work() then returns the
        # previous frame's output again (frame_decoder.cpp:47-48 returns before clearing): such frames are judged by the
        # single-frame checker below, not by the batch run
        whole = np.ones(n, bool)
        whole[cut] = False
        vm = int(((ok_g != ok_r) & whole).sum())
        pm = 0
        for f in np.nonzero(ok_g & ok_r & whole)[0]:
            if ln[f] != w_len[f] or not np.array_equal(pl[f, : ln[f]], w_pl[f, : ln[f]]):
                pm += 1
        # full comparison on a sample: every cut frame, every verdict mismatch, and 48 others
        dp = None
        sample = set(int(f) for f in cut) | set(int(f) for f in np.nonzero(ok_g != ok_r)[0]) | set(int(f) for f in rng.integers(0, n, 48))
        fm = 0
        for f in sorted(sample):
            w = ref.decode_frame(c["iq"][lts1[f]: lts1[f] + avail[f]])
            if w.n_vectors < 1:
                good = st[f] == 4
            elif not w.hdr_ok:
                good = st[f] == (1 if w.hdr_parity else 2)
            elif w.n_vectors < 1 + w.nsym:
                good = st[f] == 4 and rt[f] == w.rate and ln[f] == w.length
            else:
                if dp is None:
                    dp = dbg["depunct"].cpu().numpy()
                good = st[f] == (0 if w.crc_ok else 3) and rt[f] == w.rate and ln[f] == w.length
                good = good and np.array_equal(dp[f, : len(w.depunct)], w.depunct)
                want = w.payload if w.crc_ok else w.descrambled[2: 2 + w.length]
                good = good and bytes(pl[f, : w.length]) == bytes(want)
            fm += not good
        tot["frames"] += n
        tot["delivered"] += int(ok_r.sum())
        tot["verdict_mismatch"] += vm
        tot["payload_mismatch"] += pm
        tot["full_checked"] += len(sample)
        tot["full_mismatch"] += fm
        print("batch %2d: %5d frames, %4.1f dB, %d taps, %3d windows cut: reference delivered %5d, GPU %5d, verdict mismatches %d, "
              "payload mismatches %d, full comparison %d frames / %d mismatches"
              % (b, n, snr, taps, len(cut), int(ok_r.sum()), int(ok_g.sum()), vm, pm, len(sample), fm), flush=True)
    print(tot)
    bad = tot["verdict_mismatch"] + tot["payload_mismatch"] + tot["full_mismatch"]
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
