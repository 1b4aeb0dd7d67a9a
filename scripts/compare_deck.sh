#!/bin/bash
# GPU box: the reference's own CUDA program and graspa_b200_mc on one example deck, same seed and cycle counts: move counts
# and final energies must agree to the printed digits.
# Usage: scripts/compare_deck.sh <deck name under oracle/_ref/examples> <init cycles> <production cycles> [extra driver flags]
set -u
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
NAME="$1"; NI="$2"; NP="$3"; shift 3
OUT="$ROOT/gpurun_out/cmp_$NAME"; rm -rf "$OUT"; mkdir -p "$OUT/ref"
cp "$ROOT/oracle/_ref/examples/$NAME/"* "$OUT/ref/"; chmod u+w "$OUT/ref/"*
sed -i "s/^NumberOfInitializationCycles.*/NumberOfInitializationCycles $NI/; s/^NumberOfEquilibrationCycles.*/NumberOfEquilibrationCycles 0/; s/^NumberOfProductionCycles.*/NumberOfProductionCycles $NP/" "$OUT/ref/simulation.input"
( cd "$OUT/ref" && timeout 1500 "$ROOT/oracle/_ref/graspa_ref_cuda.x" > output.txt 2> stderr.txt; echo "reference exit $?" )
echo "--- reference ($NAME, $NI + $NP cycles)"
grep -E "Work took" "$OUT/ref/output.txt"
grep -E "Performed|Accepted" "$OUT/ref/output.txt" | grep -v "Gibbs\|CBCF\|Volume\|Special\|Single" | head -60
sed -n '/\*\*\* FINAL STAGE \*\*\*/,/Total Energy/p' "$OUT/ref/output.txt" | grep -v "^ -->\|^      \|DNN" | head -14
grep -E "Averaged Rosenbluth Weight|Averaged Henry" "$OUT/ref/output.txt" | head -4
echo "--- graspa_b200_mc"
timeout 1500 "$ROOT/graspa_b200/host/graspa_b200_mc" "$OUT/ref" --init "$NI" --equil 0 --prod "$NP" "$@" > "$OUT/ours.txt" 2>&1; echo "exit $?"
grep -E "FINAL|DRIFT|Component|Performed|Accepted|Averaged Rosenbluth Weight|Averaged Henry|Work took|moves" "$OUT/ours.txt"
