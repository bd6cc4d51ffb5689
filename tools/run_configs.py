#!/usr/bin/env python
"""BASELINE.json configs 3, 4 and 5 on their own (the same sections bench.py puts into its JSON line: bench_configs.py).

    python tools/run_configs.py [--only 3|4|5]                                   # one GPU
    python -m torch.distributed.run --nproc-per-node N ... tools/run_configs.py   # N GPUs
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_configs as bc  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", type=int, default=0)
    ap.add_argument("--c5-frames", type=int, default=131072)
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = bc.Ctx(torch, dist if world > 1 else None, dev, stream, rank, world, local)
    for k, fn in ((3, lambda: bc.config3(ctx)), (4, lambda: bc.config4(ctx)), (5, lambda: bc.config5(ctx, frames_per_gpu=a.c5_frames))):
        if a.only and a.only != k:
            continue
        out = fn()
        if rank == 0:
            print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
