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


# ------------------------------------------------------------------------------------------------
# synthetic inputs of a sharded job: a pure function of the GLOBAL insertion index
# ------------------------------------------------------------------------------------------------
DOUBLES_PER_INSERTION = 64          # 20 double3 of the random pool (10 positions + 10 orientations), 2 uniforms, 2 spare
_GAMMA, _M1, _M2 = 0x9E3779B97F4A7C15, 0xBF58476D1CE4E5B9, 0x94D049BB133111EB


def job_uniforms_numpy(first_insertion: int, count: int, seed: int) -> np.ndarray:
    """(count, 64) uniforms in [0, 1) of insertions [first, first + count) of a job: SplitMix64 of the global element index
    (counter-based: any rank generates exactly its own range, and the union over ranks is the single-process job)"""
    idx = (np.arange(first_insertion * DOUBLES_PER_INSERTION, (first_insertion + count) * DOUBLES_PER_INSERTION, dtype=np.uint64) + np.uint64(1))
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + idx * np.uint64(_GAMMA)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(_M1)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(_M2)
        z = z ^ (z >> np.uint64(31))
    return ((z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)).reshape(count, DOUBLES_PER_INSERTION)


def job_uniforms_torch(first_insertion: int, count: int, seed: int, device):
    """the same numbers generated on `device` with 64-bit integer tensor ops (wrapping multiplies, logical shifts by masking)"""
    import torch

    def s64(v):
        return v - (1 << 64) if v >= (1 << 63) else v

    def lsr(x, k):
        return (x >> k) & ((1 << (64 - k)) - 1)
    idx = torch.arange(first_insertion * DOUBLES_PER_INSERTION + 1, (first_insertion + count) * DOUBLES_PER_INSERTION + 1, dtype=torch.int64, device=device)
    z = idx * s64(_GAMMA) + s64(seed & ((1 << 64) - 1))
    z = (z ^ lsr(z, 30)) * s64(_M1)
    z = (z ^ lsr(z, 27)) * s64(_M2)
    z = z ^ lsr(z, 31)
    return (lsr(z, 11).to(torch.float64) * (1.0 / 9007199254740992.0)).reshape(count, DOUBLES_PER_INSERTION)


def split_job_uniforms(u):
    """(count, 64) -> random pool (count * 20, 3) in the packed layout of gb_widom_inputs and uniforms (count, 2)"""
    n = u.shape[0]
    return u[:, :60].reshape(n * 20, 3), u[:, 60:62]


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
