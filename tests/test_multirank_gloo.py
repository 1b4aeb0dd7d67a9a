"""N>1 path on CPU: two gloo ranks shard a Widom job into contiguous index ranges, evaluate their range (with the oracle,
which is test infrastructure), bin on the GLOBAL insertion index and all-reduce the block sums through
graspa_b200.shard -- the same helpers bench.py uses over NCCL.  The reduced sums must equal the single-process ones."""
import os
import socket
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from graspa_b200.shard import shard_range, reduce_block_sums, widom_averages
from tests.conftest import load_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def block_sums(out8, stage, first, n_total, n_blocks=5):
    """RecordRosen (data_struct.h:627-652) + widom_energy += E*W (axpy.cu:177-185) on the global insertion index"""
    s = np.zeros((n_blocks, 12))
    for i in range(out8.shape[0]):
        b = ((first + i) * n_blocks) // n_total
        w = out8[i, 0] if stage[i] == 0 else 0.0
        s[b, 0] += w; s[b, 1] += w * w; s[b, 2] += 1.0
        s[b, 3:10] += w * out8[i, 1:8]
        s[b, 10] += 1.0 if stage[i] != 0 else 0.0
    return s


def _free_port():
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def _worker(rank, world, port, outdir):
    import torch.distributed as dist
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    box, ff, s, z = load_config("A")
    ws = orc.WidomSetup(box, ff, s, int(z["comp"]), float(z["beta"]), int(z["ntrials"]), int(z["norient"]), z["sf_ads"], z["sf_fw"])
    rnd = z["widom_rnd"]; uni = z["widom_uni"]; n = uni.shape[0]
    first, count = shard_range(n, world, rank)
    out, stage, _ = orc.widom_batch(ws, rnd[first:first + count], uni[first:first + count], nthreads=1)
    total = reduce_block_sums(block_sums(out, stage, first, n))
    if rank == 0:
        np.save(os.path.join(outdir, "sums.npy"), total)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_are_contiguous_and_cover_the_job():
    for n in (0, 1, 7, 10, 1000, 10**7 + 3):
        for world in (1, 2, 3, 4, 8):
            nxt = 0
            for r in range(world):
                first, count = shard_range(n, world, r)
                assert first == nxt and count in (n // world, n // world + 1)
                nxt = first + count
            assert nxt == n
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_two_gloo_ranks_reduce_to_the_single_process_sums(oracle):
    import torch.multiprocessing as mp
    box, ff, s, z = load_config("A")
    ws = oracle.WidomSetup(box, ff, s, int(z["comp"]), float(z["beta"]), int(z["ntrials"]), int(z["norient"]), z["sf_ads"], z["sf_fw"])
    out, stage, _ = oracle.widom_batch(ws, z["widom_rnd"], z["widom_uni"], nthreads=1)
    n = out.shape[0]
    ref = block_sums(out, stage, 0, n)
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(2, _free_port(), d), nprocs=2, join=True)
        got = np.load(os.path.join(d, "sums.npy"))
    assert got.shape == ref.shape
    assert np.array_equal(got[:, 2], ref[:, 2]) and np.array_equal(got[:, 10], ref[:, 10])          # counts are exact
    assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)) < 1e-13                      # sums differ by association only
    a, b = widom_averages(got), widom_averages(ref)
    assert abs(a["mean_W"] - b["mean_W"]) <= 1e-13 * abs(b["mean_W"]) and a["count"] == n
    # a single process is its own reduction
    assert np.array_equal(reduce_block_sums(ref), ref)


def test_reference_arm_runs_on_rank_zero_only():
    """bench.py --impl reference under torchrun: every rank but 0 exits 0 without work and without output"""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
