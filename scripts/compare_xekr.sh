#!/bin/bash
# GPU box: the reference's own CUDA program and graspa_b200_mc on the XeKr-Mixture example (identity swap, two species,
# no charges, tail corrections), same seed: move counts and final energies must agree.
# Usage: scripts/compare_xekr.sh [init cycles]
set -u
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
N="${1:-20000}"
OUT="$ROOT/gpurun_out/xekr"; rm -rf "$OUT"; mkdir -p "$OUT/ref"
cp "$ROOT/oracle/_ref/examples/XeKr-Mixture/"* "$OUT/ref/"; chmod u+w "$OUT/ref/"*
sed -i "s/^NumberOfInitializationCycles.*/NumberOfInitializationCycles $N/; s/^NumberOfProductionCycles.*/NumberOfProductionCycles 0/" "$OUT/ref/simulation.input"
( cd "$OUT/ref" && timeout 900 "$ROOT/oracle/_ref/graspa_ref_cuda.x" > output.txt 2> stderr.txt; echo "reference exit $?" )
echo "--- reference"
grep -E "Work took|Fugacity Coefficient for" "$OUT/ref/output.txt"
grep -E "Performed|Accepted" "$OUT/ref/output.txt" | grep -v "Gibbs\|CBCF\|Volume\|Special\|Rotation\|Single\|Widom" | head -40
sed -n '/\*\*\* FINAL STAGE \*\*\*/,/Total Energy/p' "$OUT/ref/output.txt" | head -16
grep -A14 "ENERGY DRIFT" "$OUT/ref/output.txt" | grep "Total Energy" | head -2
echo "--- graspa_b200_mc (fused move calls)"
timeout 900 "$ROOT/graspa_b200/host/graspa_b200_mc" "$OUT/ref" --init "$N" --prod 0 > "$OUT/ours_fused.txt" 2>&1; echo "exit $?"
grep -E "FINAL|DRIFT|Component|Identity|Work took|moves" "$OUT/ours_fused.txt"
echo "--- graspa_b200_mc (stage calls)"
timeout 900 "$ROOT/graspa_b200/host/graspa_b200_mc" "$OUT/ref" --init "$N" --prod 0 --staged > "$OUT/ours_staged.txt" 2>&1; echo "exit $?"
grep -E "FINAL|DRIFT|Component|Identity|Work took|moves" "$OUT/ours_staged.txt"
