"""GPU: several devices from one process through the C ABI (include/b200rx.h: b200rx_group_*, b200rx_gather_status;
SURVEY 8e).  Runs on however many B200s the box has (a group of one on a single-GPU box - the same code path, NCCL
communicator of size 1 - and of two or more where they exist)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _group(lib, devices, max_frames, max_payload):
    from fun_ofdm_b200.rx import Limits
    lim = Limits(max_frames, max_payload)
    g = C.c_void_p()
    dev = (C.c_int * len(devices))(*devices)
    rc = lib.b200rx_group_create(dev, len(devices), C.byref(lim), C.byref(g))
    assert rc == 0, lib.b200rx_group_last_error(None).decode()
    return g


def test_group_decodes_a_batch_across_devices_and_gathers_status(ref):
    import torch
    from fun_ofdm_b200 import rx as rxmod, tx
    lib = rxmod.load_library()
    n_dev = min(torch.cuda.device_count(), 4)
    assert n_dev >= 1
    devices = list(range(n_dev))
    rng = np.random.default_rng(99)
    n = 96 * n_dev
    rates = rng.integers(0, 11, n).astype(np.uint8)
    payloads = [rng.integers(0, 256, int(rng.integers(1, 700)), dtype=np.uint8).tobytes() for _ in range(n)]
    corpus = tx.build_corpus(payloads, rates, snr_db=24.0, lead_in=16, seed=5, threads=4)
    g = _group(lib, devices, 128, 1500)
    try:
        assert lib.b200rx_group_size(g) == n_dev
        first = np.zeros(n_dev + 1, np.uint32)
        lib.b200rx_group_plan(g, corpus["avail"].ctypes.data_as(C.c_void_p), n, first.ctypes.data_as(C.c_void_p))
        assert first[0] == 0 and first[-1] == n and np.all(np.diff(first.astype(np.int64)) > 0)
        # host buffers: one call, every device decodes its shard
        iq = np.ascontiguousarray(corpus["iq"])
        payload = np.zeros((n, 1500), np.uint8)
        length = np.zeros(n, np.uint16)
        rate = np.zeros(n, np.uint8)
        status = np.full(n, 77, np.uint8)
        rc = lib.b200rx_group_decode_batch(g, iq.ctypes.data_as(C.c_void_p), len(iq), corpus["lts1"].ctypes.data_as(C.c_void_p),
                                           corpus["avail"].ctypes.data_as(C.c_void_p), n, payload.ctypes.data_as(C.c_void_p), 1500,
                                           length.ctypes.data_as(C.c_void_p), rate.ctypes.data_as(C.c_void_p),
                                           status.ctypes.data_as(C.c_void_p))
        assert rc == 0, lib.b200rx_group_last_error(g).decode()
        # the same frames through a single handle on device 0 and through the reference
        single = rxmod.Receiver(0, n, 1500)
        p1, l1, r1, s1 = single.decode_batch(corpus["iq"], corpus["lts1"], corpus["avail"])
        single.close()
        assert np.array_equal(status, s1) and np.array_equal(length, l1) and np.array_equal(rate, r1)
        n_ok = 0
        for f in range(n):
            off, m = int(corpus["lts1"][f]), int(corpus["avail"][f])
            want = ref.decode_frame(corpus["iq"][off: off + m])
            assert (status[f] == 0) == bool(want.hdr_ok and want.crc_ok), f
            if status[f] == 0:
                assert bytes(payload[f, : length[f]]) == bytes(want.payload) == payloads[f], f
                n_ok += 1
        assert n_ok >= n // 2

        # device buffers + NCCL gather: equal shards, per-device arrays
        per = n // n_dev
        keep = []
        arr = lambda ptrs: (C.c_void_p * n_dev)(*ptrs)  # noqa: E731
        iq_p, l_p, a_p, pl_p, ln_p, rt_p, st_p, ga_p, ns, nf = [], [], [], [], [], [], [], [], [], []
        for d in devices:
            dev = torch.device("cuda:%d" % d)
            f0 = d * per
            lo = int(corpus["lts1"][f0])
            hi = int(corpus["lts1"][f0 + per - 1] + corpus["avail"][f0 + per - 1])
            t_iq = torch.from_numpy(iq[lo:hi].view(np.float64)).to(dev)
            t_l = torch.from_numpy((corpus["lts1"][f0: f0 + per] - lo).astype(np.uint64).view(np.int64)).to(dev)
            t_a = torch.from_numpy(corpus["avail"][f0: f0 + per].astype(np.uint32).view(np.int32)).to(dev)
            t_pl = torch.zeros((per, 1500), dtype=torch.uint8, device=dev)
            t_ln = torch.zeros(per, dtype=torch.int16, device=dev)
            t_rt = torch.zeros(per, dtype=torch.uint8, device=dev)
            t_st = torch.full((per,), 99, dtype=torch.uint8, device=dev)
            t_ga = torch.full((n_dev * per,), 98, dtype=torch.uint8, device=dev)
            torch.cuda.synchronize(dev)
            keep += [t_iq, t_l, t_a, t_pl, t_ln, t_rt, t_st, t_ga]
            iq_p.append(t_iq.data_ptr()); l_p.append(t_l.data_ptr()); a_p.append(t_a.data_ptr()); pl_p.append(t_pl.data_ptr())
            ln_p.append(t_ln.data_ptr()); rt_p.append(t_rt.data_ptr()); st_p.append(t_st.data_ptr()); ga_p.append(t_ga.data_ptr())
            ns.append(hi - lo); nf.append(per)
        rc = lib.b200rx_group_decode_batch_dev(g, arr(iq_p), (C.c_uint64 * n_dev)(*ns), arr(l_p), arr(a_p), (C.c_uint32 * n_dev)(*nf),
                                               arr(pl_p), 1500, arr(ln_p), arr(rt_p), arr(st_p))
        assert rc == 0, lib.b200rx_group_last_error(g).decode()
        sums = np.zeros(4, np.uint64)
        rc = lib.b200rx_gather_status(g, arr(st_p), per, arr(ga_p), sums.ctypes.data_as(C.c_void_p))
        assert rc == 0, lib.b200rx_group_last_error(g).decode()
        for d in devices:  # every device holds everybody's status bytes
            gathered = keep[8 * d + 7].cpu().numpy()
            assert np.array_equal(gathered, status[: n_dev * per]), d
        ok = int((status[: n_dev * per] == 0).sum())
        assert int(sums[0]) == ok and int(sums[0] + sums[1]) == n_dev * per
        assert int(sums[2]) == int(length[: n_dev * per][status[: n_dev * per] == 0].astype(np.int64).sum())
        assert lib.b200rx_group_synchronize(g) == 0
    finally:
        lib.b200rx_group_destroy(g)
