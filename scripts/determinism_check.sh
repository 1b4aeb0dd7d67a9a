#!/bin/bash
# GPU box: run-to-run determinism of the host driver on a deck (default: the two-box Gibbs example), in three modes run as concurrent
# pairs (time slicing perturbs the timing): one-kernel moves, stage calls, one-kernel moves with every stream drained after each move.
# Usage: scripts/determinism_check.sh [deck] [init cycles] [production cycles]
D="oracle/_ref/examples/${1:-NVT-Gibbs}"; NI="${2:-60}"; NP="${3:-60}"
mkdir -p gpurun_out/det
run() { env $3 ./graspa_b200/host/graspa_b200_mc $D --init $NI --prod $NP $2 --trace gpurun_out/det/$1.txt > gpurun_out/det/$1.out 2>&1; }
run fused_a "" "X=1" & run fused_b "" "X=1" & run staged_a "--staged" "X=1" & run staged_b "--staged" "X=1" &
run sync_a "" "GB_SYNC_EVERY_MOVE=1" & run sync_b "" "GB_SYNC_EVERY_MOVE=1" &
wait
for m in fused staged sync; do
  if cmp -s gpurun_out/det/${m}_a.txt gpurun_out/det/${m}_b.txt; then echo "$m: the two runs are identical ($(wc -l < gpurun_out/det/${m}_a.txt) moves)";
  else echo "$m: runs DIFFER: $(cmp gpurun_out/det/${m}_a.txt gpurun_out/det/${m}_b.txt)"; fi
done
cmp -s gpurun_out/det/fused_a.txt gpurun_out/det/staged_a.txt && echo "fused_a == staged_a"; cmp -s gpurun_out/det/sync_a.txt gpurun_out/det/staged_a.txt && echo "sync_a == staged_a"
rm -f gpurun_out/det/*.txt
