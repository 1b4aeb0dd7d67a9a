#!/bin/bash
# GPU box: the reference's own CUDA program and graspa_b200_mc on the CO2_NaX_Zeolite example (charged framework, 55 movable
# Na+ as a separated framework component, block pockets, Peng-Robinson fugacity), same seed.
# Usage: scripts/compare_nax.sh [init cycles] [production cycles]
set -u
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
NI="${1:-2000}"; NP="${2:-0}"
OUT="$ROOT/gpurun_out/nax"; rm -rf "$OUT"; mkdir -p "$OUT/ref"
cp "$ROOT/oracle/_ref/examples/CO2_NaX_Zeolite/"* "$OUT/ref/"; chmod u+w "$OUT/ref/"*
sed -i "s/^NumberOfInitializationCycles.*/NumberOfInitializationCycles $NI/; s/^NumberOfProductionCycles.*/NumberOfProductionCycles $NP/" "$OUT/ref/simulation.input"
( cd "$OUT/ref" && timeout 900 "$ROOT/oracle/_ref/graspa_ref_cuda.x" > output.txt 2> stderr.txt; echo "reference exit $?" )
echo "--- reference"
grep -E "Work took|Fugacity Coefficient for|Replicated block" "$OUT/ref/output.txt"
grep -E "Performed|Accepted" "$OUT/ref/output.txt" | grep -v "Gibbs\|CBCF\|Volume\|Special\|Single\|Widom\|Identity" | head -40
sed -n '/\*\*\* FINAL STAGE \*\*\*/,/Total Energy/p' "$OUT/ref/output.txt" | grep -v "^ -->\|^      " | head -16
sed -n '/\*\*\* RUNNING DELTA_E (FINAL - CREATE MOLECULE) \*\*\*/,/Total Energy/p' "$OUT/ref/output.txt" | grep -v "^ -->\|^      " | head -16
echo "--- graspa_b200_mc (fused move calls)"
timeout 900 "$ROOT/graspa_b200/host/graspa_b200_mc" "$OUT/ref" --init "$NI" --prod "$NP" > "$OUT/ours_fused.txt" 2>&1; echo "exit $?"
grep -E "INITIAL|FINAL|RUNNING|DRIFT|Component|Translation|Rotation|Work took|moves" "$OUT/ours_fused.txt"
echo "--- graspa_b200_mc (stage calls)"
timeout 900 "$ROOT/graspa_b200/host/graspa_b200_mc" "$OUT/ref" --init "$NI" --prod "$NP" --staged > "$OUT/ours_staged.txt" 2>&1; echo "exit $?"
grep -E "FINAL|RUNNING|DRIFT|Component|Translation|Rotation|Work took|moves" "$OUT/ours_staged.txt"
