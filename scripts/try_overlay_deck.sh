#!/bin/bash
# The reference bound to libgraspa_b200.so on one example deck, beside the stock trace build.  Run on the GPU box:
#   DECK=XeKr-Mixture CYCLES=2000 scripts/try_overlay_deck.sh  -> gpurun_out/overlay_<deck>/
set -u
ROOT=$(cd "$(dirname "$0")/.." && pwd)
DECK=${DECK:-XeKr-Mixture}
OUT=$ROOT/gpurun_out/overlay_$DECK; mkdir -p $OUT
D=$(mktemp -d); cp -r $ROOT/oracle/_ref/examples/$DECK/* $D/; chmod -R u+w $D
sed -i -e "s/^NumberOfInitializationCycles.*/NumberOfInitializationCycles ${CYCLES:-2000}/" -e "s/^NumberOfProductionCycles.*/NumberOfProductionCycles ${PROD:-0}/" -e "s/^NumberOfEquilibrationCycles.*/NumberOfEquilibrationCycles 0/" $D/simulation.input
for exe in graspa_ref_cuda_trace.x graspa_ref_overlay.x; do
  (cd $D && GRASPA_TRACE=$OUT/$exe.trace timeout 600 ${WRAP:-} $ROOT/oracle/_ref/$exe > $OUT/$exe.stdout 2> $OUT/$exe.stderr; echo "rc=$?" >> $OUT/$exe.stdout)
done
tail -c 600 $OUT/graspa_ref_overlay.x.stdout; tail -15 $OUT/graspa_ref_overlay.x.stderr
wc -l $OUT/*.trace
cmp $OUT/graspa_ref_cuda_trace.x.trace $OUT/graspa_ref_overlay.x.trace | head -2
