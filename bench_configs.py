"""BASELINE.json configs[2..4] as sections of the bench line (bench.py imports this; tools/run_configs.py prints them alone).

    config 3  rate sweep: every rate (8 of 802.11a + the reference's 3 extra 2/3 rates), 1500-byte frames
    config 4  Viterbi only: K=7 64-state decode of 1/2, 2/3, 3/4 punctured inputs, 12 096- and 32 832-step trellises
    config 5  mixed-length frames (64-4095 B) at 54 Mbps through a 4-tap multipath channel, genie tags

Every section runs on all ranks at once (one rank per GPU, weak scaling: the per-GPU work is fixed), is timed with CUDA
events on the launching stream between barriers, takes the max over ranks, and sums the counts over ranks; there is no
data-path collective.  Inputs are generated in HBM (device-side frame builder + channel, torch for the coded symbols of
config 4).  Each section also carries a parity sample: rank 0 copies a few frames back and compares status, LENGTH and
bytes with the checker (oracle/, the unmodified reference compiled under oracle/_ref) - outside the timed regions,
as the checker, never as the thing measured.
"""
import numpy as np

RATE_NAMES = ["1/2 BPSK", "2/3 BPSK", "3/4 BPSK", "1/2 QPSK", "2/3 QPSK", "3/4 QPSK",
              "1/2 QAM16", "2/3 QAM16", "3/4 QAM16", "2/3 QAM64", "3/4 QAM64"]
FAIL_SNR = [3, 5, 7, 6, 8, 10, 12, 14, 16, 20, 22]  # dB at which roughly half of the 1500-byte frames fail, per rate


class Ctx:
    """What every section needs: torch device / stream of this rank, the distributed helpers, the checker (rank 0)."""

    def __init__(self, torch, dist, dev, stream, rank, world, local_rank, tuning=None):
        self.torch, self.dist, self.dev, self.stream = torch, dist, dev, stream
        self.rank, self.world, self.local_rank = rank, world, local_rank
        self.tuning = tuning or {}
        self._checker = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, ms, counts):
        """max over ranks of the time, sum over ranks of the counts"""
        t = self.torch
        tm = t.tensor([ms], dtype=t.float64, device=self.dev)
        tc = t.tensor([int(c) for c in counts], dtype=t.int64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(tm, op=self.dist.ReduceOp.MAX)
            self.dist.all_reduce(tc)
        return float(tm.item()), [int(x) for x in tc.tolist()]

    def checker(self):
        """TEST INFRASTRUCTURE (oracle/): only used to compare a sample of results, on rank 0, outside timed regions."""
        if self._checker is None:
            from oracle import bind
            self._checker = (bind.ref(), "reference") if bind.have_ref() else (bind.port(), "port")
        return self._checker

    def receiver(self, max_frames, max_payload):
        import fun_ofdm_b200 as fo
        rx = fo.Receiver(self.local_rank, max_frames, max_payload)
        for k, v in self.tuning.items():
            rx.set_tuning(k, v)
        rx.set_stream(self.stream.cuda_stream)
        return rx

    def timed(self, fn, warmup, passes):
        """ms per pass of fn() on self.stream, barrier + synchronize on both sides"""
        t = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for _ in range(passes):
            fn()
        e1.record(self.stream)
        self.barrier()
        return e0.elapsed_time(e1) / passes


def _outs(t, dev, n, stride):
    return dict(payload=t.zeros((n, stride), dtype=t.uint8, device=dev), length=t.zeros(n, dtype=t.int16, device=dev),
                rate=t.zeros(n, dtype=t.uint8, device=dev), status=t.zeros(n, dtype=t.uint8, device=dev))


def _check_frames(ctx, corpus, out, picks):
    """Frames `picks` of a device corpus against the checker: (checked, mismatches)."""
    chk, _ = ctx.checker()
    st = out["status"].cpu().numpy()
    ln = out["length"].cpu().numpy().astype(np.uint16).astype(np.int64)
    lts1 = corpus["lts1"].cpu().numpy()
    avail = corpus["avail"].cpu().numpy()
    bad = 0
    for f in picks:
        w0, m = int(lts1[f]), int(avail[f])
        win = corpus["iq"][2 * w0: 2 * (w0 + m)].cpu().numpy().view(np.complex128)
        w = chk.decode_frame(win)
        want_st = 0 if (w.hdr_ok and w.crc_ok) else (3 if w.hdr_ok else (1 if w.hdr_parity else 2))
        if w.hdr_ok and w.n_vectors < 1 + w.nsym:
            want_st = 4
        good = int(st[f]) == want_st
        if good and w.hdr_ok and want_st in (0, 3):
            want = w.payload if w.crc_ok else w.descrambled[2: 2 + w.length]
            got = out["payload"][f, : w.length].cpu().numpy()
            good = bytes(got) == bytes(want) and int(ln[f]) == w.length
        bad += not good
    return len(picks), bad


def config3(ctx, frames=2048, payload=1500, passes=8, check=4, depth=4):
    """Rate sweep.  Throughput point: 30 dB, `depth` batches in flight.  Parity sample: `check` frames at 30 dB and `check`
    at the failing SNR."""
    from fun_ofdm_b200 import tx
    t = ctx.torch
    rx = ctx.receiver(frames, payload)
    rx.set_pipeline_depth(depth)
    points, checked, bad = [], 0, 0
    for rate in range(11):
        rng = np.random.default_rng(1000 + 16 * ctx.rank + rate)
        pl = rng.integers(0, 256, (frames, payload), dtype=np.uint8)
        c = tx.build_corpus_dev(pl, np.full(frames, rate, np.uint8), snr_db=30.0, seed=300 + rate + 64 * ctx.rank,
                                device=ctx.local_rank, stream=ctx.stream.cuda_stream)
        outs = [_outs(t, ctx.dev, frames, payload) for _ in range(depth)]
        k = [0]

        def go():
            o = outs[k[0] % depth]
            k[0] += 1
            rx.decode_batch_dev(c["iq"], c["lts1"], c["avail"], o["payload"], o["length"], o["rate"], o["status"])

        def run():
            for _ in range(depth):
                go()
            rx.join(0)

        ms = ctx.timed(run, 1, max(1, passes // depth)) / depth
        o = outs[0]
        ok = (o["status"] == 0)
        n_ok = int(ok.sum())
        same = bool(t.equal(o["payload"][ok], t.from_numpy(pl).to(ctx.dev)[ok]))
        same = same and all(bool(t.equal(x["status"], o["status"])) for x in outs[1:])
        if ctx.rank == 0 and check:
            a, b = _check_frames(ctx, c, o, list(range(0, frames, max(1, frames // check)))[:check])
            checked, bad = checked + a, bad + b
            cf = tx.build_corpus_dev(pl[:check], np.full(check, rate, np.uint8), snr_db=float(FAIL_SNR[rate]), seed=900 + rate,
                                     device=ctx.local_rank, stream=ctx.stream.cuda_stream)
            of = _outs(t, ctx.dev, check, payload)
            rx.decode_batch_dev(cf["iq"], cf["lts1"], cf["avail"], of["payload"], of["length"], of["rate"], of["status"])
            rx.synchronize()
            a, b = _check_frames(ctx, cf, of, list(range(check)))
            checked, bad = checked + a, bad + b
        ms, (n_all, ok_all, diff) = ctx.reduce(ms, [frames, n_ok, 0 if same else 1])
        points.append({"rate": RATE_NAMES[rate], "frames": n_all, "frames_ok": ok_all, "ms_per_batch": ms,
                       "payload_mbit_s": ok_all * payload * 8 / ms / 1e3, "frames_per_s": n_all / ms * 1e3,
                       "ok_payloads_equal_transmitted": diff == 0})
        del c, outs
    rx.close()
    return {"workload": "BASELINE configs[2]: rate sweep, %d x %d-byte frames per GPU and rate, AWGN 30 dB, %d batches in flight; "
                        "parity sample at 30 dB and at the SNR where about half the frames fail" % (frames, payload, depth),
            "n_gpus": ctx.world, "points": points, "parity_sample_frames": checked, "parity_mismatches": bad}


def _coded_symbols(ctx, n_distinct, nb, punc, sigma, seed):
    """Config 4 input on the device: random information bits -> viterbi::conv_encode (viterbi.cpp:39-62: sr = (sr << 1) | bit,
    outputs parity(sr & 121), parity(sr & 91), for nb + 6 bits read from the data) -> puncturer::puncture -> hard 0 / 255 +
    Gaussian(sigma), clamped -> puncturer::depuncture (erasures 127).  Returns (uint8 [n_distinct, 2 * (nb + 6)], bits)."""
    t = ctx.torch
    g = t.Generator(device=ctx.dev)
    g.manual_seed(seed)
    steps = nb + 6
    bits = t.randint(0, 2, (n_distinct, steps), generator=g, device=ctx.dev, dtype=t.int32)
    bits[:, nb:] = 0   # a zero tail, so that the traceback's start state 0 is the true one (the reference's ppdu does not: Q6)
    pad = t.zeros((n_distinct, 6), dtype=t.int32, device=ctx.dev)
    b = t.cat([pad, bits], 1)  # b[:, 6 + t] = bit t; earlier bits are 0 (sr starts at 0)

    def tap(k):  # bit from k steps ago
        return b[:, 6 - k: 6 - k + steps]

    c0 = tap(0) ^ tap(3) ^ tap(4) ^ tap(5) ^ tap(6)    # 121 = 0b1111001
    c1 = tap(0) ^ tap(1) ^ tap(3) ^ tap(4) ^ tap(6)    # 91  = 0b1011011
    coded = t.stack([c0, c1], 2).reshape(n_distinct, 2 * steps).to(t.float32) * 255.0
    noisy = t.clamp(t.round(coded + sigma * t.randn(coded.shape, generator=g, device=ctx.dev)), 0, 255).to(t.uint8)
    # puncture + depuncture = overwrite the punctured positions with the erasure value (puncturer.cpp:41-63, 94-118)
    idx = t.arange(2 * steps, device=ctx.dev)
    if punc == 1:      # 2/3: of every 4 coded bits keep {0, 2, 3}
        erase = (idx % 4) == 1
    elif punc == 2:    # 3/4: of every 6 coded bits keep {0, 1, 3, 5}
        erase = ((idx % 6) == 2) | ((idx % 6) == 4)
    else:
        erase = t.zeros_like(idx, dtype=t.bool)
    noisy[:, erase] = 127
    return noisy, bits


def config4(ctx, frames=16384, passes=5, check=8):
    """Viterbi only (b200rx_viterbi_batch_dev = viterbi::conv_decode for a batch)."""
    t = ctx.torch
    points, checked, bad = [], 0, 0
    for steps in (12096, 32832):
        nb = steps - 6
        max_payload = 1500 if steps == 12096 else 4095
        n = frames if steps == 12096 else max(256, frames // 4)  # same number of trellis steps per launch, roughly
        rx = ctx.receiver(n, max_payload)
        rx.set_pipeline_depth(1)
        for punc, name in ((0, "1/2"), (1, "2/3"), (2, "3/4")):
            base = 64
            sym, bits = _coded_symbols(ctx, base, nb, punc, 40.0, 4000 + 10 * punc + ctx.rank)
            d_sym = sym.repeat(n // base, 1).contiguous()
            d_bits = t.full((n,), nb, dtype=t.int32, device=ctx.dev)
            d_out = t.zeros((n, (nb + 7) // 8), dtype=t.uint8, device=ctx.dev)
            ms = ctx.timed(lambda: rx.viterbi_batch_dev(d_sym, d_bits, nb, d_out), 2, passes)
            st = rx.stats()
            # decoded bytes MSB first against the transmitted bits (sigma 40: error-free), and a sample against the checker
            got = d_out[:base]
            want = (bits[:, : nb // 8 * 8].reshape(base, -1, 8) * (2 ** t.arange(7, -1, -1, device=ctx.dev, dtype=t.int32))).sum(2).to(t.uint8)
            wrong = int((got[:, : want.shape[1]] != want).any(1).sum())
            if ctx.rank == 0 and check:
                chk, _ = ctx.checker()
                hs, ho = sym[:check].cpu().numpy(), d_out[:check].cpu().numpy()
                for f in range(check):
                    w = chk.conv_decode(hs[f], nb)
                    checked += 1
                    bad += not np.array_equal(ho[f, : len(w)], w)
            ms, (n_all, wrong_all) = ctx.reduce(ms, [n, wrong])
            acs_ms, _ = ctx.reduce(st["viterbi_ms"], [0])
            points.append({"code_rate": name, "trellis_steps": steps, "frames": n_all, "ms": ms,
                           "decoded_gbit_s": n_all * nb / ms / 1e6, "acs_per_s": 64.0 * steps * n_all / (acs_ms * 1e-3),
                           "acs_kernel_ms": acs_ms, "distinct_frames_differing_from_transmitted_bits": wrong_all})
            del d_sym, d_out
        rx.close()
    return {"workload": "BASELINE configs[3]: Viterbi-only K=7 64-state decode, 1/2, 2/3, 3/4 punctured inputs (hard 0/255 + "
                        "Gaussian sigma 40, erasures 127), %d frames per GPU of 12 096 steps and %d of 32 832 steps"
                        % (frames, max(256, frames // 4)),
            "n_gpus": ctx.world, "points": points, "parity_sample_frames": checked, "parity_mismatches": bad}


def config5(ctx, frames_per_gpu=131072, sub=8192, passes=2, check=24, depth=3):
    """Mixed-length frames through multipath, sharded by Viterbi work; 2^20 frames on 8 GPUs = 131 072 per GPU."""
    from fun_ofdm_b200 import shard, tx
    t = ctx.torch
    total = frames_per_gpu * ctx.world
    rng = np.random.default_rng(55)
    lengths_all = rng.integers(64, 4096, total)
    order = np.argsort(-lengths_all, kind="stable")       # alike frames share an ACS warp
    work = shard.trellis_steps(np.full(total, 10), lengths_all[order])
    bounds = shard.balanced_ranges(work, ctx.world)
    mine = order[bounds[ctx.rank]: bounds[ctx.rank + 1]]
    rx = ctx.receiver(sub, 4095)
    rx.set_pipeline_depth(depth)
    batches = []
    for s in range(0, len(mine), sub):
        idx = mine[s: s + sub]
        prng = np.random.default_rng(1_000_003 * (ctx.rank + 1) + s)
        blob = prng.integers(0, 256, int(lengths_all[idx].sum()), dtype=np.uint8)
        offs = np.concatenate([[0], np.cumsum(lengths_all[idx])])
        payloads = [blob[offs[i]: offs[i + 1]] for i in range(len(idx))]
        c = tx.build_corpus_dev(payloads, np.full(len(idx), 10, np.uint8), snr_db=30.0, multipath_taps=4, seed=9000 + int(idx[0]),
                                device=ctx.local_rank, stream=ctx.stream.cuda_stream)
        c.update(payloads=payloads, n=len(idx), out=_outs(t, ctx.dev, len(idx), 4095))
        batches.append(c)

    def decode_all():
        for c in batches:
            o = c["out"]
            rx.decode_batch_dev(c["iq"], c["lts1"], c["avail"], o["payload"], o["length"], o["rate"], o["status"])
        rx.join(0)

    ms = ctx.timed(decode_all, 1, passes)
    ok_frames = ok_bits = wrong = checked = bad = 0
    sample_bytes = 0
    per_batch = max(1, check // max(1, len(batches)))
    for c in batches:
        o = c["out"]
        st = o["status"].cpu().numpy()
        ln = o["length"].cpu().numpy().astype(np.uint16).astype(np.int64)
        ok = st == 0
        ok_frames += int(ok.sum())
        ok_bits += int(ln[ok].sum()) * 8
        sample_bytes += c["iq"].numel() * 8
        pl = o["payload"].cpu().numpy()
        for f in np.nonzero(ok)[0][::17]:  # every 17th CRC-OK payload against what was transmitted
            wrong += bytes(pl[f, : ln[f]]) != bytes(c["payloads"][f])
        if ctx.rank == 0 and check:
            fails = list(np.nonzero(~ok)[0][: per_batch // 2])
            picks = fails + list(np.linspace(0, c["n"] - 1, max(1, per_batch - len(fails))).astype(int))
            a, b = _check_frames(ctx, c, o, picks)
            checked, bad = checked + a, bad + b
    rx.close()
    ms, (n_all, ok_all, bits_all, wrong_all, steps_all, bytes_all) = ctx.reduce(
        ms, [len(mine), ok_frames, ok_bits, wrong, int(work[bounds[ctx.rank]: bounds[ctx.rank + 1]].sum()), sample_bytes])
    return {"workload": "BASELINE configs[4]: %d mixed-length frames (64-4095 B) per GPU (2^20 on 8 GPUs) at 54 Mbps, 4-tap "
                        "multipath + 30 dB AWGN, genie tags; sorted by length, sharded by Viterbi work, generated in HBM, decoded "
                        "in sub-batches of %d with %d in flight" % (frames_per_gpu, sub, depth),
            "n_gpus": ctx.world, "frames": n_all, "frames_ok": ok_all, "ms_per_pass": ms, "frames_per_s": n_all / ms * 1e3,
            "crc_ok_payload_mbit_s": bits_all / ms / 1e3, "trellis_steps": steps_all, "sample_bytes": bytes_all,
            "ok_payloads_differing_from_transmitted": wrong_all, "parity_sample_frames": checked, "parity_mismatches": bad}
