"""Multi-GPU plumbing of the Widom path.

Ghost insertions never change the system, so a job of `n_total` insertions is cut into contiguous index ranges, one per
rank (one process per GPU); every rank holds a replica of the system and evaluates its range with `gb_widom_batch`
(bins assigned on the GLOBAL insertion index through `gb_widom_inputs.global_first/global_n`).  The only exchange is one
all-reduce of the `n_blocks x 12` block sums, mirroring the reference's `RosenbluthWeight` accumulators
(data_struct.h:491-499, axpy.cu:177-185).  GCMC inside one box does not shard (sequential Markov chain): replicas only.
"""
from __future__ import annotations

import numpy as np


def shard_range(n_total: int, world: int, rank: int):
    """contiguous range [first, first + count) of rank `rank`; counts differ by at most one, earlier ranks get the extras"""
    if world < 1 or not (0 <= rank < world) or n_total < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(n_total, world)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def reduce_block_sums(sums: np.ndarray, device=None):
    """element-wise SUM of the (n_blocks, 12) block sums over all ranks (NCCL when `device` is a CUDA device, gloo on CPU).
    A single-process run returns its input."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return np.array(sums, dtype=np.float64, copy=True)
    t = torch.from_numpy(np.ascontiguousarray(sums, dtype=np.float64)).clone()
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def widom_averages(sums: np.ndarray):
    """per-block <W>, its standard error over blocks, and the failed fraction, from reduced block sums
    (columns: sumW, sumW2, count, 7 x sum(W*E), n_failed, reserved)"""
    sums = np.asarray(sums, dtype=np.float64)
    cnt = np.maximum(sums[:, 2], 1.0)
    w_block = sums[:, 0] / cnt
    total = float(sums[:, 2].sum())
    mean = float(sums[:, 0].sum() / max(total, 1.0))
    err = float(w_block.std(ddof=1) / np.sqrt(len(w_block))) if len(w_block) > 1 else 0.0
    return dict(mean_W=mean, block_W=w_block, stderr_W=err, failed_fraction=float(sums[:, 10].sum() / max(total, 1.0)), count=total)
