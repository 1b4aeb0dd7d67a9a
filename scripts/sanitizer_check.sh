#!/bin/bash
# GPU box: compute-sanitizer over the round-2 kernels: memcheck and racecheck of the cell-sorted Widom stage (parity tests of both
# pair stages and of the replay), and of the resident move server on short GCMC runs (Xe/Kr with identity swaps, CO2-MFI with Ewald).
cd "$(dirname "$0")/.."
S=/usr/local/cuda/bin/compute-sanitizer
run() { echo "== $*"; timeout 900 "$@" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error|hazard|cycles_per_s" | cut -c1-200 | tail -6; }
run $S --tool memcheck python -m pytest tests/test_gpu_parity.py -x -q -k "widom_batch_vs_golden or resumes or large_trial_batches"
run $S --tool racecheck python -m pytest tests/test_gpu_parity.py -x -q -k "widom_batch_vs_golden and cells"
for deck in "XeKr-Mixture 400" "CO2-MFI 60"; do
  set -- $deck
  D=$(mktemp -d /tmp/san.XXXX); cp -r oracle/_ref/examples/$1/* $D/; chmod -R u+w $D
  GB_MOVE_SERVER_IDLE_MS=60000 GB_MOVE_TIMEOUT_MS=600000 run $S --tool memcheck ./graspa_b200/host/graspa_b200_mc $D --init $2 --equil 0 --prod 0
  GB_MOVE_SERVER_IDLE_MS=60000 GB_MOVE_TIMEOUT_MS=600000 run $S --tool racecheck ./graspa_b200/host/graspa_b200_mc $D --init $2 --equil 0 --prod 0
  rm -rf $D
done
