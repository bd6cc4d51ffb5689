"""CPU tests of the multi-rank host logic (world_size 2, gloo): contiguous sharding, balancing by
trellis steps, and the status gather / counter reduction that is the path's only collective."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")


def test_ranges():
    from fun_ofdm_b200 import shard
    assert shard.even_ranges(4096, 8) == [512 * r for r in range(9)]
    assert shard.even_ranges(10, 4) == [0, 2, 5, 7, 10]
    rng = np.random.default_rng(0)
    lengths = rng.integers(64, 4096, 5000)
    rates = rng.choice([0, 2, 3, 5, 6, 8, 9, 10], 5000)
    work = shard.trellis_steps(rates, lengths)
    assert work[0] == -(-(16 + 8 * (lengths[0] + 4) + 6) // [24, 32, 36, 48, 64, 72, 96, 128, 144, 192, 216][rates[0]]) * \
        [24, 32, 36, 48, 64, 72, 96, 128, 144, 192, 216][rates[0]]
    b = shard.balanced_ranges(work, 8)
    assert b[0] == 0 and b[-1] == 5000 and all(x <= y for x, y in zip(b, b[1:]))
    per = [int(work[b[r]: b[r + 1]].sum()) for r in range(8)]
    assert max(per) - min(per) <= 2 * int(work.max())
    assert shard.balanced_ranges(np.array([5]), 4) == [0, 0, 0, 1, 1] or shard.balanced_ranges(np.array([5]), 4)[-1] == 1


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from fun_ofdm_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_local = 6
    bounds = shard.even_ranges(n_local * world, world)
    status = torch.tensor([(bounds[rank] + i) % 5 for i in range(n_local)], dtype=torch.uint8)
    counters = torch.tensor([int((status == 0).sum()), int((status != 0).sum()), 100 * (rank + 1), 7], dtype=torch.int64)
    all_status, counters = shard.gather_status(status, counters, world)
    q.put((rank, all_status.tolist(), counters.tolist()))
    dist.destroy_process_group()


def test_status_gather_world2():
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want_status = [i % 5 for i in range(12)]
    for rank, st, cnt in res:
        assert st == want_status
        assert cnt == [sum(1 for x in want_status if x == 0), sum(1 for x in want_status if x != 0), 300, 14]
