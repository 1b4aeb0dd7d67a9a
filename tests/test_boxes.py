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
