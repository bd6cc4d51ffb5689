"""Small invocations of EVERY kernel of libb200rx.so, for compute-sanitizer (memcheck / racecheck / synccheck).

    compute-sanitizer --tool memcheck  python tools/sanitize_run.py
    compute-sanitizer --tool racecheck python tools/sanitize_run.py

Sizes are chosen so that every kernel template instance that the library can launch runs at least once while the
sanitizer's 10-100x slowdown stays within minutes: all 11 rates, short and ragged frames, the three sample formats,
the rotation path of the raw-capture entry points, the pull kernel (pinned caller buffer, > 512 ordered frames), the
header-only and Viterbi-only entry points, the device-side generator, pipeline lanes and host slots.
Results are checked against what was transmitted (the sanitizer run is not a parity test; tests/ is).
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fun_ofdm_b200 as fo  # noqa: E402
from fun_ofdm_b200 import tx  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(5)
    lib = fo.load_library()

    # ---- host-buffer decode, all rates, ragged lengths (K1 / K2 LB=4 / K3), then the three formats ----
    rates = list(range(11)) * 2
    payloads = [rng.integers(0, 256, int(rng.integers(0, 300)), dtype=np.uint8).tobytes() for _ in rates]
    corpus = tx.build_corpus(payloads, rates, snr_db=30.0, lead_in=16, seed=3, threads=2)
    rx = fo.Receiver(0, max_frames=640, max_payload_bytes=400)
    payload, length, rate, status = rx.decode_batch(corpus["iq"], corpus["lts1"], corpus["avail"])
    ok = sum(1 for f in range(len(rates)) if status[f] == 0 and bytes(payload[f, : length[f]]) == payloads[f])
    print("decode_batch fc64: %d/%d ok" % (ok, len(rates)))
    for fmt, conv in ((fo.FMT_FC32, lambda x: x.astype(np.complex64)),
                      (fo.FMT_SC16, None)):
        if fmt == fo.FMT_SC16:
            scale = float(np.max(np.abs(corpus["iq"].view(np.float64)))) / 30000.0
            wire = np.clip(np.rint(corpus["iq"].view(np.float64) / scale), -32768, 32767).astype(np.int16)
            rx.set_sample_format(fmt, scale)
        else:
            wire = conv(corpus["iq"])
            rx.set_sample_format(fmt)
        p2, l2, r2, s2 = rx.decode_batch(wire, corpus["lts1"], corpus["avail"])
        print("decode_batch fmt %d: %d ok" % (fmt, int((s2 == 0).sum())))
    rx.set_sample_format(fo.FMT_FC64)

    # ---- header-only entry point ----
    lts1 = np.ascontiguousarray(corpus["lts1"], np.uint64)
    av = np.ascontiguousarray(corpus["avail"], np.uint32)
    iq = np.ascontiguousarray(corpus["iq"])
    hl = np.zeros(len(rates), np.uint16)
    hr = np.zeros(len(rates), np.uint8)
    hs = np.zeros(len(rates), np.uint8)
    rc = lib.b200rx_decode_headers(rx.h, iq.ctypes.data, len(iq), lts1.ctypes.data, av.ctypes.data, len(rates),
                                   hl.ctypes.data, hr.ctypes.data, hs.ctypes.data)
    assert rc == 0
    print("decode_headers: %d ok" % int((hs == 0).sum()))

    # ---- pinned caller buffer, > 512 ordered frames: chunked pipeline + pull kernel, two calls in flight ----
    n = 600
    rates2 = [int(r) for r in rng.integers(0, 11, n)]
    payloads2 = [rng.integers(0, 256, int(rng.integers(1, 40)), dtype=np.uint8).tobytes() for _ in range(n)]
    c2 = tx.build_corpus(payloads2, rates2, snr_db=30.0, lead_in=8, seed=4, threads=4)
    iq2 = np.ascontiguousarray(c2["iq"])
    sizes = [iq2.nbytes, n * 8, n * 4] + [n * 400, n * 2, n, n] * 2
    ptrs = []
    for sz in sizes:
        p = C.c_void_p()
        assert lib.b200rx_host_alloc(C.byref(p), sz) == 0
        ptrs.append(p)
    C.memmove(ptrs[0], iq2.ctypes.data, iq2.nbytes)
    l64 = c2["lts1"].astype(np.uint64)
    a32 = c2["avail"].astype(np.uint32)
    C.memmove(ptrs[1], l64.ctypes.data, n * 8)
    C.memmove(ptrs[2], a32.ctypes.data, n * 4)
    t1 = rx.submit_batch_ptr(ptrs[0], len(iq2), ptrs[1], ptrs[2], n, ptrs[3], 400, ptrs[4], ptrs[5], ptrs[6])
    t2 = rx.submit_batch_ptr(ptrs[0], len(iq2), ptrs[1], ptrs[2], n, ptrs[7], 400, ptrs[8], ptrs[9], ptrs[10])
    rx.wait(t1)
    rx.wait(t2)
    st = np.ctypeslib.as_array(C.cast(ptrs[6], C.POINTER(C.c_uint8)), shape=(n,)).copy()
    st_b = np.ctypeslib.as_array(C.cast(ptrs[10], C.POINTER(C.c_uint8)), shape=(n,)).copy()
    assert np.array_equal(st, st_b)
    print("submit_batch (pull kernel, 2 in flight): %d/%d ok" % (int((st == 0).sum()), n))
    for p in ptrs:
        lib.b200rx_host_free(p)

    # ---- device-resident decode over pipeline lanes (K2 LB=3 needs >= 444 frames), debug taps ----
    d_iq = torch.from_numpy(iq2.view(np.float64)).to(dev)
    d_l = torch.from_numpy(c2["lts1"].astype(np.int64)).to(dev)
    d_a = torch.from_numpy(c2["avail"].astype(np.int32)).to(dev)
    rx.set_pipeline_depth(2)
    outs = [dict(payload=torch.zeros((n, 400), dtype=torch.uint8, device=dev), length=torch.zeros(n, dtype=torch.int16, device=dev),
                 rate=torch.zeros(n, dtype=torch.uint8, device=dev), status=torch.zeros(n, dtype=torch.uint8, device=dev))
            for _ in range(2)]
    for j in range(4):
        o = outs[j % 2]
        rx.decode_batch_dev(d_iq, d_l, d_a, o["payload"], o["length"], o["rate"], o["status"])
    rx.join(0)
    rx.synchronize()
    assert torch.equal(outs[0]["status"], outs[1]["status"])
    assert np.array_equal(outs[0]["status"].cpu().numpy(), st)
    rx.set_pipeline_depth(1)
    dbg = dict(equalized=torch.zeros((n, 4, 48, 2), dtype=torch.float64, device=dev),
               decoded=torch.zeros((n, 64), dtype=torch.uint8, device=dev),
               header_field=torch.zeros(n, dtype=torch.int32, device=dev),
               depunct=torch.zeros((n, 1024), dtype=torch.uint8, device=dev))
    o = outs[0]
    rx.decode_batch_dev(d_iq, d_l, d_a, o["payload"], o["length"], o["rate"], o["status"], debug=dbg)
    rx.synchronize()
    print("decode_batch_dev lanes + debug taps: %d ok" % int((o["status"] == 0).sum()))

    # ---- Viterbi-only entry point (bm_from_symbols_kernel, raw-mode traceback) ----
    nb = 24 * 20 - 6
    sym = torch.from_numpy(rng.integers(0, 256, (32, 2 * (nb + 6)), dtype=np.uint8)).to(dev)
    bits = torch.full((32,), nb, dtype=torch.int32, device=dev)
    out = torch.zeros((32, (nb + 7) // 8), dtype=torch.uint8, device=dev)
    rx.viterbi_batch_dev(sym, bits, nb, out)
    rx.synchronize()
    print("viterbi_batch_dev: done")

    # ---- device generator + raw capture: detector, scan, LTS sync, frame list, rotated front end ----
    cap = tx.build_corpus_dev(payloads[:8], rates[:8], snr_db=28.0, lead_in=400, seed=11, multipath_taps=2)
    x = np.concatenate([cap["iq"].cpu().numpy().view(np.complex128), np.zeros(2048, complex)])
    got, info = rx.receive(x)
    print("receive: %d frames found, %d payloads" % (info["n_frames"], len(got)))
    d_x = torch.from_numpy(x.view(np.float64)).to(dev)
    tags = torch.zeros(len(x), dtype=torch.uint8, device=dev)
    res = rx.sync_dev(d_x, tags=tags)
    print("sync_dev: %s" % res)
    rx.close()

    # ---- two-phase passes through the host adapters (pack_pass_kernel, tag_find / tag_frames kernels, select mask,
    # tagged sample format), a group of one device with the NCCL gather ----
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_chain import Chain
    from test_gpu_block import Block
    xs = np.concatenate([x, x * 0.5, np.zeros(4096, complex)])
    ch = Chain(max_frames=64, max_payload=400, depth=3, max_lag=2)
    got_chain, _, _ = ch.run(xs, 4096, max_out=64, stride=400)
    ch.close()
    print("b200_receiver_chain (passes): %d payloads" % len(got_chain))
    tg = np.zeros(len(corpus["iq"]), np.uint8)
    tg[corpus["lts1"].astype(np.int64)] = 4
    blk = Block(max_frames=64, max_payload=400, depth=3, max_lag=2)
    got_blk, _, _ = blk.run(corpus["iq"], tg, 5000, max_out=64, stride=400)
    blk.close()
    print("b200_rx (tagged passes): %d payloads of %d frames" % (len(got_blk), len(rates)))
    from fun_ofdm_b200.rx import Limits
    g = C.c_void_p()
    lim = Limits(64, 400)
    assert lib.b200rx_group_create((C.c_int * 1)(0), 1, C.byref(lim), C.byref(g)) == 0
    gp = np.zeros((len(rates), 400), np.uint8)
    gl = np.zeros(len(rates), np.uint16)
    gr = np.zeros(len(rates), np.uint8)
    gs = np.zeros(len(rates), np.uint8)
    assert lib.b200rx_group_decode_batch(g, iq.ctypes.data, len(iq), lts1.ctypes.data, av.ctypes.data, len(rates), gp.ctypes.data,
                                         400, gl.ctypes.data, gr.ctypes.data, gs.ctypes.data) == 0
    lib.b200rx_group_destroy(g)
    print("group of one device: %d ok" % int((gs == 0).sum()))
    print("sanitize_run: all entry points exercised")


if __name__ == "__main__":
    main()
