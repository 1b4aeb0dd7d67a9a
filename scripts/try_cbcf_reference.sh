#!/bin/bash
# Does the STOCK reference program run a CB/CFC deck at all?  CO2-MFI with CBCFProbability added (no example deck of the reference
# uses CB/CFC).  Run on the GPU box: scripts/try_cbcf_reference.sh  -> gpurun_out/cbcf_ref/
set -u
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/gpurun_out/cbcf_ref; mkdir -p $OUT
D=$(mktemp -d); cp $ROOT/oracle/_ref/examples/CO2-MFI/* $D/; chmod u+w $D/*
sed -i -e 's/^NumberOfInitializationCycles.*/NumberOfInitializationCycles 3000/' -e 's/CreateNumberOfMolecules  0/CreateNumberOfMolecules  8/' \
       -e 's/^\( *\)SwapProbability\(.*\)$/\1SwapProbability\2\n\1CBCFProbability          1.0\n\1LambdaType ShiMaginn/' $D/simulation.input
cat $D/simulation.input > $OUT/simulation.input
for exe in graspa_ref_cuda.x graspa_ref_cuda_trace.x; do
  (cd $D && timeout 300 $ROOT/oracle/_ref/$exe > $OUT/$exe.stdout 2> $OUT/$exe.stderr; echo "rc=$?" >> $OUT/$exe.stdout)
  [ -f $D/output.txt ] && cp $D/output.txt $OUT/$exe.output.txt
  ls $D > $OUT/$exe.files
done
tail -5 $OUT/graspa_ref_cuda.x.stdout
grep -n "CBCF\|DRIFT\|Lambda" $OUT/graspa_ref_cuda.x.stdout | tail -40
