#!/bin/bash
# GPU box: move-by-move parity sweep over the example decks the host driver supports (scripts/compare_trace.sh each).
set -u
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
N="${1:-2000}"
for deck in CO2-MFI XeKr-Mixture CO2_NaX_Zeolite Ar_MgMOF74_UFF Bae-Mixture BlockPocket CO2_MgMOF74_UFF Tail-Correction Ionic-MOF-mixtures TIP4PEW-MgMOF-LJ1264 Restart-Examples Restart-LAMMPS; do
  echo "=== $deck ($N initialisation cycles)"
  bash "$ROOT/scripts/compare_trace.sh" "$deck" "$N" 0 2>&1 | tail -6
  grep -E "DRIFT|Work took" "$ROOT/gpurun_out/trace_$deck/ours.txt" | head -2
  grep -E "Work took" "$ROOT/gpurun_out/trace_$deck/ref/output.txt" | head -1
  tail -2 "$ROOT/gpurun_out/trace_$deck/ours.txt" | grep -v "^{" | cut -c1-300
done
