"""Independent boxes across GPUs (graspa_b200/boxes.py): the deal of boxes to GPUs and the gathering of per-box results.
The orchestration is exercised on CPU with a stand-in driver script; the real driver needs a GPU (tests/test_gpu_boxes)."""
import json
import os
import stat

import pytest

from graspa_b200.boxes import assign_boxes, run_boxes


def test_boxes_are_dealt_round_robin():
    assert assign_boxes(5, 2) == [[0, 2, 4], [1, 3]]
    assert assign_boxes(3, 8) == [[0], [1], [2], [], [], [], [], []]
    assert assign_boxes(0, 2) == [[], []]
    q = assign_boxes(17, 8)
    assert sorted(b for g in q for b in g) == list(range(17)) and max(len(g) for g in q) - min(len(g) for g in q) <= 1
    with pytest.raises(ValueError):
        assign_boxes(3, 0)


def test_results_are_gathered_per_box_in_input_order(tmp_path):
    fake = tmp_path / "fake_driver.sh"
    fake.write_text("""#!/bin/bash
# stand-in for graspa_b200_mc: echoes the flags it was given in the driver's output format; the GPU of a box is the ONE
# device its process can see (CUDA_VISIBLE_DEVICES), addressed as device 0
dev=-1; p=0
while [ $# -gt 0 ]; do case "$1" in --device) [ "$2" = 0 ] && dev=$CUDA_VISIBLE_DEVICES; shift;; --pressure) p=$2; shift;; esac; shift; done
echo "FINAL   VDW [Host-Host]: 0.0, Total: -$p"
echo "ENERGY DRIFT (FINAL - INITIAL - RUNNING) Total Energy: 1.0e-10"
echo "{\\"pressure_pa\\": $p, \\"temperature\\": 300, \\"loading\\": [{\\"component\\": \\"X\\", \\"molecules\\": $dev, \\"production_average\\": 1.5}]}"
echo "{\\"moves\\": 10, \\"cycles\\": 10, \\"seconds\\": 0.1, \\"moves_per_s\\": 100, \\"cycles_per_s\\": 100}"
""")
    os.chmod(fake, os.stat(fake).st_mode | stat.S_IEXEC)
    pts = [{"pressure": 1e4 * (k + 1)} for k in range(5)]
    res, wall = run_boxes("deck", pts, gpus=2, init=1, prod=1, driver=str(fake))
    assert [r["box"] for r in res] == [0, 1, 2, 3, 4]
    assert [r["gpu"] for r in res] == [0, 1, 0, 1, 0]
    for k, r in enumerate(res):
        assert r["returncode"] == 0 and abs(r["pressure_pa"] - 1e4 * (k + 1)) < 1e-6
        assert r["loading"][0]["molecules"] == r["gpu"] and r["run"]["cycles"] == 10
        assert abs(r["final_total_energy"] + 1e4 * (k + 1)) < 1e-6 and r["energy_drift"] == 1.0e-10


def test_missing_driver_fails_loudly(tmp_path):
    with pytest.raises(FileNotFoundError):
        run_boxes("deck", [{}], driver=str(tmp_path / "nope"))


def test_a_rank_runs_its_box_on_its_own_device(tmp_path):
    """bench.py's multibox_xekr: every rank of a torchrun launch calls run_boxes with ONE point and devices=[LOCAL_RANK]"""
    fake = tmp_path / "fake_driver.sh"
    fake.write_text("""#!/bin/bash
echo "{\\"loading\\": [{\\"component\\": \\"X\\", \\"molecules\\": $CUDA_VISIBLE_DEVICES, \\"production_average\\": 0}]}"
echo "{\\"moves\\": 10, \\"cycles\\": 10, \\"seconds\\": 0.1, \\"moves_per_s\\": 100, \\"cycles_per_s\\": 100}"
""")
    os.chmod(fake, os.stat(fake).st_mode | stat.S_IEXEC)
    env_before = os.environ.pop("CUDA_VISIBLE_DEVICES", None)
    try:
        for local in (0, 3, 7):
            res, _ = run_boxes("deck", [{"pressure": 1e5}], gpus=1, init=1, prod=0, driver=str(fake), devices=[local])
            assert res[0]["loading"][0]["molecules"] == local
        os.environ["CUDA_VISIBLE_DEVICES"] = "4,5,6"          # a restricted launch: local rank 1 owns physical device 5
        res, _ = run_boxes("deck", [{"pressure": 1e5}], gpus=1, init=1, prod=0, driver=str(fake), devices=[1])
        assert res[0]["loading"][0]["molecules"] == 5
    finally:
        os.environ.pop("CUDA_VISIBLE_DEVICES", None)
        if env_before is not None:
            os.environ["CUDA_VISIBLE_DEVICES"] = env_before


def test_multibox_summary_of_the_bench_line():
    """aggregate = all boxes' cycles / the slowest box's Monte Carlo loop time; a failed box withholds the aggregate instead of flattering it"""
    import bench
    recs = [{"rank": r, "pressure_pa": p, "returncode": 0, "cycles": 1000, "mc_seconds": t, "cycles_per_s": 1000 / t, "process_seconds": t + 2.0, "energy_drift": d}
            for r, (p, t, d) in enumerate([(1e4, 0.010, 1e-11), (3e4, 0.012, -3e-10), (1e5, 0.008, None)])]
    s = bench.multibox_summary(recs)
    assert s["boxes"] == 3 and abs(s["value"] - 3000 / 0.012) < 1e-6
    assert abs(s["value_incl_process_start"] - 3000 / 2.012) < 1e-6
    assert abs(s["speedup_vs_one_after_the_other"] - 0.030 / 0.012) < 1e-9 and s["max_abs_energy_drift"] == 3e-10
    bad = bench.multibox_summary(recs[:2] + [{"rank": 2, "error": "driver missing"}])
    assert "value" not in bad and bad["boxes"] == 3
