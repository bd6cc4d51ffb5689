"""What the box can feed its GPUs: bare pinned host->device copies on 1 .. N GPUs at once (one rank per GPU, torchrun),
nothing decoded.  `e2e` of bench.py moves 320 MB of std::complex<double> samples per 4096-frame step per GPU; this is the
ceiling that number lives under.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/h2d_ceiling.py

Rank 0 prints one JSON object: for k = 1, 2, 4, .. N concurrently copying ranks, the aggregate GB/s (max over ranks of the
wall time around `reps` copies of `mbytes` MB, barrier on both sides) and what that would allow in decoded payload Gbit/s at
BASELINE config 2 (78 096 sample bytes per 12 000 payload bits)."""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mbytes, reps = 320, 20
    host = torch.empty(mbytes * 1000 * 1000, dtype=torch.uint8).pin_memory()
    host.fill_(rank + 1)
    dev = torch.empty_like(host, device="cuda")
    back = torch.empty(6 * 1000 * 1000, dtype=torch.uint8).pin_memory()
    points = []
    k = 1
    ks = []
    while k < world:
        ks.append(k)
        k *= 2
    ks.append(world)
    for k in ks:
        active = rank < k
        for _ in range(3):
            if active:
                dev.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        if active:
            for _ in range(reps):
                dev.copy_(host, non_blocking=True)
                back.copy_(dev[: back.numel()], non_blocking=True)  # the results of a step: 6 MB device -> host
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0 if active else 0.0], device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            dist.barrier()
        dt = float(dt.item())
        gbs = k * reps * host.numel() / dt / 1e9
        points.append({"gpus_copying": k, "aggregate_h2d_gbytes_per_s": gbs, "per_gpu_gbytes_per_s": gbs / k,
                       "ms_per_320MB_step": 1e3 * dt / reps,
                       "payload_gbit_s_this_allows_at_config2_fc64": gbs * 1e9 / (4880 * 16 + 1504) * 12000 / 1e9})
    if rank == 0:
        print(json.dumps({"what": "bare pinned cudaMemcpyAsync host->device, %d MB per copy, %d copies, one rank per GPU" % (mbytes, reps),
                          "host_cpus": os.cpu_count(), "points": points}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
