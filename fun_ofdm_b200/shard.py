"""Multi-GPU plumbing of the hot path: frames are independent, so a batch is cut into contiguous
per-rank ranges with NO data-path collective; the only communication is the final gather of the
per-frame status bytes and the reduction of the counters (NCCL on GPUs; gloo in the CPU tests)."""
import numpy as np

from .rx import RATE_PARAMS, num_symbols


def even_ranges(n_frames, world):
    """Contiguous equal split: rank r owns [bounds[r], bounds[r+1])."""
    return [(n_frames * r) // world for r in range(world + 1)]


def trellis_steps(rates, lengths):
    """Viterbi work of each frame: nsym * dbps trellis steps (the dominant cost of the path)."""
    return np.array([num_symbols(int(r), int(l)) * RATE_PARAMS[int(r)][2] for r, l in zip(rates, lengths)],
                    dtype=np.int64)


def balanced_ranges(work, world):
    """Contiguous split with (nearly) equal total work per rank (mixed-length corpora, BASELINE config 5).
    Greedy on the prefix sum: boundary r is the first frame at which the running work reaches r/world."""
    work = np.asarray(work, dtype=np.int64)
    total = int(work.sum())
    prefix = np.concatenate([[0], np.cumsum(work)])
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        b = int(np.searchsorted(prefix, target, side="left"))
        bounds.append(max(min(b, len(work)), bounds[-1]))
    bounds.append(len(work))
    return bounds


def gather_status(status, counters, world):
    """All-gather of the per-frame status bytes (equal shard sizes) and all-reduce of the counters.
    `status`: uint8 tensor [n_local]; `counters`: int64 tensor.  Returns (status_all [world*n_local], counters)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return status, counters
    out = torch.empty(status.numel() * world, dtype=status.dtype, device=status.device)
    dist.all_gather_into_tensor(out, status)
    dist.all_reduce(counters)
    return out, counters
