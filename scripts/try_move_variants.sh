#!/bin/bash
# kernel experiments on the one-kernel moves: GCMC throughput of the example decks with alternative builds of the library
# (directories holding a libgraspa_b200.so; "default" = the in-tree one).  Usage: scripts/try_move_variants.sh default build_dbg/x ...
set -u
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
for v in "$@"; do
  echo "=== $v"
  for run in "XeKr-Mixture 40000" "CO2-MFI 4000" "CO2_NaX_Zeolite 20000"; do
    set -- $run
    if [ "$v" = default ]; then L=""; else L="$ROOT/$v"; fi
    for rep in 1 2; do
      LD_LIBRARY_PATH="$L" "$ROOT/graspa_b200/host/graspa_b200_mc" "$ROOT/oracle/_ref/examples/$1" --init "$2" --prod 0 2>&1 | grep -E '"cycles_per_s"|^FINAL' | sed -e 's/.*"cycles_per_s": \([0-9.]*\).*/  '"$1"' cycles\/s \1/' -e 's/^FINAL.*Total: /  total /'
    done
  done
done
