#!/bin/bash
# The reference bound to libgraspa_b200.so (oracle/_ref/graspa_ref_overlay.x) on the CO2-MFI deck with CBCFProbability added.
# Run on the GPU box -> gpurun_out/cbcf_overlay/
set -u
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/gpurun_out/cbcf_overlay; mkdir -p $OUT
D=$(mktemp -d); cp $ROOT/oracle/_ref/examples/CO2-MFI/* $D/; chmod u+w $D/*
sed -i -e "s/^NumberOfInitializationCycles.*/NumberOfInitializationCycles ${CYCLES:-3000}/" -e 's/CreateNumberOfMolecules  0/CreateNumberOfMolecules  8/' \
       -e 's/^\( *\)SwapProbability\(.*\)$/\1SwapProbability\2\n\1CBCFProbability          1.0\n\1LambdaType ShiMaginn/' $D/simulation.input
for exe in graspa_ref_cuda_trace.x graspa_ref_overlay.x; do
  (cd $D && CUDA_LAUNCH_BLOCKING=${BLOCKING:-0} GRASPA_TRACE=$OUT/$exe.trace timeout 600 ${WRAP:-} $ROOT/oracle/_ref/$exe > $OUT/$exe.stdout 2> $OUT/$exe.stderr; echo "rc=$?" >> $OUT/$exe.stdout)
done
tail -c 1500 $OUT/graspa_ref_overlay.x.stdout; tail -5 $OUT/graspa_ref_overlay.x.stderr
wc -l $OUT/*.trace
cmp $OUT/graspa_ref_cuda_trace.x.trace $OUT/graspa_ref_overlay.x.trace | head -2
