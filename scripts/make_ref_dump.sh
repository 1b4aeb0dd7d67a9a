#!/bin/bash
# GPU box: runs the instrumented reference program (oracle/build_ref.sh dump) on three example decks and leaves the dumps in
# gpurun_out/ref_dump_<deck>.txt.  tests/golden/make_ref_dump.py (build container) turns them into committed fixtures.
set -u
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
X="$ROOT/oracle/_ref/graspa_ref_cuda_dump.x"
mkdir -p "$ROOT/gpurun_out"
run() {  # deck init prod ndump
  local W="$(mktemp -d /tmp/refdump.XXXXXX)"
  cp -r "$ROOT/oracle/_ref/examples/$1/"* "$W/"; chmod -R u+w "$W"
  sed -i "s/^NumberOfInitializationCycles.*/NumberOfInitializationCycles $2/; s/^NumberOfEquilibrationCycles.*/NumberOfEquilibrationCycles 0/; s/^NumberOfProductionCycles.*/NumberOfProductionCycles $3/" "$W/simulation.input"
  ( cd "$W" && GRASPA_DUMP="$ROOT/gpurun_out/ref_dump_$1.txt" GRASPA_DUMP_N="$4" timeout 600 "$X" > "$ROOT/gpurun_out/ref_dump_$1.out" 2>&1; echo "$1 exit $?" )
  wc -l "$ROOT/gpurun_out/ref_dump_$1.txt"
  rm -rf "$W"
}
run Henrys_coefficient 0 400 300
run CO2-MFI 60 0 200
run CO2_NaX_Zeolite 3000 0 150
run BlockPocket 1500 0 150
