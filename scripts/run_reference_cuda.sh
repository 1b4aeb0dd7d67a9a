#!/bin/bash
# Runs the reference's own CUDA build (oracle/_ref/graspa_ref_cuda.x, built by oracle/build_ref.sh) on its example
# decks on the GPU box: the "reference CUDA build on the same B200" baseline of BASELINE.md 2a and the ground truth
# for accept/reject-sequence and <W> parity.  Usage: scripts/run_reference_cuda.sh <outdir> [henry_cycles] [mfi_cycles]
set -u
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
OUT="${1:-$ROOT/gpurun_out/ref_runs}"
HENRY="${2:-20000}"
MFI="${3:-5000}"
BIN="$ROOT/oracle/_ref/graspa_ref_cuda.x"
mkdir -p "$OUT"
run() { # name, sed expression
  local name="$1"; shift
  local d="$OUT/$name"; rm -rf "$d"; mkdir -p "$d"; cp "$ROOT/oracle/_ref/examples/$name/"* "$d/"; chmod u+w "$d"/*
  for e in "$@"; do sed -i "$e" "$d/simulation.input"; done
  ( cd "$d"; t0=$(date +%s.%N); timeout 900 "$BIN" > output.txt 2> stderr.txt; rc=$?; t1=$(date +%s.%N)
    echo "exit $rc" >> stderr.txt; echo "wall_seconds $(echo "$t1 - $t0" | bc -l 2>/dev/null || python3 -c "print($t1-$t0)")" >> stderr.txt )
  grep -E "Work took|wall_seconds|exit" "$d/output.txt" "$d/stderr.txt" | sed "s|$OUT/||"
}
run Henrys_coefficient "s/^NumberOfProductionCycles.*/NumberOfProductionCycles $HENRY/"
run CO2-MFI "s/^NumberOfInitializationCycles.*/NumberOfInitializationCycles $MFI/"
run XeKr-Mixture "s/^NumberOfInitializationCycles.*/NumberOfInitializationCycles 20000/" "s/^NumberOfProductionCycles.*/NumberOfProductionCycles 0/"
run CO2_NaX_Zeolite
