#!/usr/bin/env python
"""BASELINE.json configs 3, 4 and 5 on one B200 (configs[1] is bench.py's job): throughput with device-resident
input (CUDA events, 3 warm-up + 10 timed passes) and bit-exact parity against the checker on a sample.
Prints one JSON object per config; `gpurun` sessions tee it into gpurun_out/ and it is committed under profiles/.

    python tools/run_configs.py [--frames5 16384]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fun_ofdm_b200 as fo  # noqa: E402
from fun_ofdm_b200 import shard, tx  # noqa: E402
from oracle import bind  # noqa: E402

DEV = torch.device("cuda:0")
NAMES = bind.RATE_NAMES


def checker():
    return bind.ref() if bind.have_ref() else bind.port()


def to_dev(c):
    return (torch.from_numpy(c["iq"].view(np.float64)).to(DEV), torch.from_numpy(c["lts1"].astype(np.int64)).to(DEV),
            torch.from_numpy(c["avail"].astype(np.int32)).to(DEV))


def timed_decode(rx, d_iq, d_l, d_a, n, stride, passes=10):
    payload = torch.zeros((n, stride), dtype=torch.uint8, device=DEV)
    length = torch.zeros(n, dtype=torch.int16, device=DEV)
    rate = torch.zeros(n, dtype=torch.uint8, device=DEV)
    status = torch.zeros(n, dtype=torch.uint8, device=DEV)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    rx.set_stream(stream.cuda_stream)
    for _ in range(3):
        rx.decode_batch_dev(d_iq, d_l, d_a, payload, length, rate, status)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(passes):
        rx.decode_batch_dev(d_iq, d_l, d_a, payload, length, rate, status)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / passes
    return ms, payload.cpu().numpy(), length.cpu().numpy().astype(np.uint16).astype(int), rate.cpu().numpy(), status.cpu().numpy()


def parity_sample(c, payload, length, status, idx):
    chk = checker()
    bad = 0
    for f in idx:
        off, m = int(c["lts1"][f]), int(c["avail"][f])
        w = chk.decode_frame(c["iq"][off: off + m])
        st = 0 if (w.hdr_ok and w.crc_ok) else (3 if w.hdr_ok else (1 if w.hdr_parity else 2))
        if w.hdr_ok and w.n_vectors < 1 + w.nsym:
            st = 4
        ok = int(status[f]) == st
        if ok and w.hdr_ok and st in (0, 3):
            want = w.payload if w.crc_ok else w.descrambled[2: 2 + w.length]
            ok = bytes(payload[f, : w.length]) == bytes(want) and int(length[f]) == w.length
        bad += not ok
    return bad


def config3(n=512):
    """Rate sweep: all 11 rates, 1500-byte frames, at a comfortable SNR and at one where many frames fail."""
    out = []
    rx = fo.Receiver(0, n, 1500)
    for rate in range(11):
        for snr in (30.0, [3, 5, 7, 6, 8, 10, 12, 14, 16, 20, 22][rate]):
            rng = np.random.default_rng(1000 + rate)
            pl = rng.integers(0, 256, (n, 1500), dtype=np.uint8)
            c = tx.build_corpus(pl, np.full(n, rate, np.uint8), snr_db=snr, seed=300 + rate)
            ms, payload, length, r, status = timed_decode(rx, *to_dev(c), n, 1500)
            ok = status == 0
            assert np.array_equal(payload[ok], pl[ok])
            bad = parity_sample(c, payload, length, status, range(0, n, 16))
            out.append({"rate": NAMES[rate], "snr_db": snr, "frames": n, "frames_ok": int(ok.sum()), "ms": ms,
                        "payload_mbit_s": int(ok.sum()) * 12000 / ms / 1e3, "frames_per_s": n / ms * 1e3,
                        "parity_sample": n // 16, "parity_mismatches": bad})
    rx.close()
    return {"config": "3: rate sweep, 1500-byte frames, 512 frames per point, bit-exact vs reference on a 32-frame sample",
            "points": out}


def config4(n=4096):
    """Viterbi-only: conv_encode -> puncture -> hard 0/255 + Gaussian -> clamp -> depuncture(127) -> b200rx_viterbi_batch_dev."""
    ref = checker()
    out = []
    rx = fo.Receiver(0, n, 1500)
    steps = 12096
    nb = steps - 6
    for rate, name in ((0, "1/2"), (1, "2/3"), (2, "3/4")):
        for sigma in (10, 40, 60, 90):
            rng = np.random.default_rng(40 + rate)
            base = 64  # distinct frames, tiled to n
            sym = np.zeros((base, 2 * rx.max_steps), np.uint8)
            want = []
            for i in range(base):
                data = rng.integers(0, 256, (nb + 13) // 8 + 1, dtype=np.uint8)
                coded = ref.conv_encode(data, nb)
                txs = ref.puncture(coded, rate).astype(np.float64) * 255.0
                rxs = np.clip(np.rint(txs + sigma * rng.standard_normal(len(txs))), 0, 255).astype(np.uint8)
                dep = ref.depuncture(rxs, rate)
                sym[i, : len(dep)] = dep
                want.append(ref.conv_decode(dep, nb))
            d_sym = torch.from_numpy(np.tile(sym, (n // base, 1))).to(DEV)
            d_bits = torch.full((n,), nb, dtype=torch.int32, device=DEV)
            d_out = torch.zeros((n, rx.max_steps // 8 + 8), dtype=torch.uint8, device=DEV)
            stream = torch.cuda.Stream()
            torch.cuda.set_stream(stream)
            rx.set_stream(stream.cuda_stream)
            for _ in range(3):
                rx.viterbi_batch_dev(d_sym, d_bits, nb, d_out)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(stream)
            for _ in range(10):
                rx.viterbi_batch_dev(d_sym, d_bits, nb, d_out)
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            o = d_out.cpu().numpy()
            bad = sum(not np.array_equal(o[f, : len(want[f % base])], want[f % base]) for f in range(n))
            st = rx.stats()
            out.append({"code_rate": name, "sigma": sigma, "frames": n, "trellis_steps": steps, "ms": ms,
                        "decoded_gbit_s": n * nb / ms / 1e6, "acs_per_s": 64.0 * steps * n / (st["viterbi_ms"] * 1e-3),
                        "acs_kernel_ms": st["viterbi_ms"], "mismatching_frames_vs_reference": int(bad)})
    rx.close()
    return {"config": "4: Viterbi-only K=7 64-state, 4096 frames x 12 096 steps, every frame compared with viterbi::conv_decode",
            "points": out}


def config5(n_total=16384, sub=4096):
    """Mixed-length frames (64..4095 B) at 54 Mbps through 4-tap multipath + 30 dB AWGN, genie tags; frames sorted by
    length so that the four frames sharing an ACS warp are alike; processed in sub-batches of `sub` frames."""
    rng = np.random.default_rng(55)
    lengths = rng.integers(64, 4096, n_total)
    order = np.argsort(-lengths, kind="stable")
    rx = fo.Receiver(0, sub, 4095)
    tot_ms, ok_frames, ok_bits, bad, checked = 0.0, 0, 0, 0, 0
    work = shard.trellis_steps(np.full(n_total, 10), lengths)
    for s in range(0, n_total, sub):
        idx = order[s: s + sub]
        payloads = [rng.integers(0, 256, int(lengths[i]), dtype=np.uint8).tobytes() for i in idx]
        c = tx.build_corpus(payloads, np.full(len(idx), 10, np.uint8), snr_db=30.0, multipath_taps=4, seed=9000 + s)
        ms, payload, length, r, status = timed_decode(rx, *to_dev(c), len(idx), 4095, passes=3)
        tot_ms += ms
        ok = status == 0
        ok_frames += int(ok.sum())
        ok_bits += int(length[ok].sum()) * 8
        for f in np.nonzero(ok)[0][:64]:
            assert bytes(payload[f, : length[f]]) == payloads[f]
        sample = list(range(0, len(idx), max(1, len(idx) // 24)))
        bad += parity_sample(c, payload, length, status, sample)
        checked += len(sample)
    rx.close()
    return {"config": "5: %d mixed-length frames (64-4095 B) at 54 Mbps, 4-tap multipath + 30 dB AWGN, genie tags, one GPU, "
                      "sub-batches of %d sorted by length" % (n_total, sub),
            "frames": n_total, "frames_ok": ok_frames, "ms_total": tot_ms, "frames_per_s": n_total / tot_ms * 1e3,
            "crc_ok_payload_mbit_s": ok_bits / tot_ms / 1e3, "trellis_steps": int(work.sum()),
            "parity_sample": checked, "parity_mismatches": bad}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames5", type=int, default=16384)
    ap.add_argument("--only", type=int, default=0)
    a = ap.parse_args()
    for k, fn in ((3, config3), (4, config4), (5, lambda: config5(a.frames5))):
        if a.only and a.only != k:
            continue
        print(json.dumps(fn()), flush=True)


if __name__ == "__main__":
    main()
