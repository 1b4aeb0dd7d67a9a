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


def _box_worker(rank, world, port, outdir, fake_driver):
    """what bench.py does for its multibox_xekr key: every rank runs ITS isotherm point on ITS device, the records are gathered with
    all_gather_object and rank 0 summarises them"""
    import json
    import torch.distributed as dist
    import bench
    from graspa_b200.boxes import run_boxes
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    os.environ.pop("CUDA_VISIBLE_DEVICES", None)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res, _ = run_boxes("deck", [{"pressure": bench.MULTIBOX_PRESSURES[rank]}], gpus=1, init=100, prod=0, driver=fake_driver, devices=[rank])
    r = res[0]; run = r["run"]
    mine = {"rank": rank, "pressure_pa": r["point"]["pressure"], "returncode": r["returncode"], "cycles": run["cycles"], "mc_seconds": run["seconds"],
            "cycles_per_s": run["cycles_per_s"], "process_seconds": r["seconds"], "device_seen": r["loading"][0]["molecules"], "energy_drift": r.get("energy_drift")}
    allrec = [None] * world
    dist.all_gather_object(allrec, mine)
    if rank == 0:
        json.dump(bench.multibox_summary(allrec), open(os.path.join(outdir, "multibox.json"), "w"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gloo_ranks_run_one_box_each_and_rank_zero_gathers_them(tmp_path):
    import json
    import stat
    import torch.multiprocessing as mp
    fake = tmp_path / "fake_driver.sh"
    fake.write_text("""#!/bin/bash
# stand-in for graspa_b200_mc: a box at higher pressure takes longer; reports the ONE device it can see
p=0; while [ $# -gt 0 ]; do case "$1" in --pressure) p=$2; shift;; esac; shift; done
secs=$(python3 -c "print(0.001 * (1 + ($p > 20000)))")
echo "ENERGY DRIFT (FINAL - INITIAL - RUNNING) Total Energy: 2.0e-11"
echo "{\\"pressure_pa\\": $p, \\"loading\\": [{\\"component\\": \\"Xe\\", \\"molecules\\": $CUDA_VISIBLE_DEVICES, \\"production_average\\": 0}]}"
echo "{\\"moves\\": 100, \\"cycles\\": 100, \\"seconds\\": $secs, \\"moves_per_s\\": 1, \\"cycles_per_s\\": 1}"
""")
    os.chmod(fake, os.stat(fake).st_mode | stat.S_IEXEC)
    mp.spawn(_box_worker, args=(2, _free_port(), str(tmp_path), str(fake)), nprocs=2, join=True)
    s = json.load(open(tmp_path / "multibox.json"))
    assert s["boxes"] == 2 and [r["rank"] for r in s["per_box"]] == [0, 1]
    assert [r["device_seen"] for r in s["per_box"]] == [0, 1]                      # every rank's box ran on its own device
    assert [r["pressure_pa"] for r in s["per_box"]] == [1e4, 3e4]
    assert abs(s["value"] - 200 / 0.002) < 1e-6                                    # all cycles / the slowest box's loop time
    assert abs(s["speedup_vs_one_after_the_other"] - 1.5) < 1e-9 and s["max_abs_energy_drift"] == 2.0e-11
