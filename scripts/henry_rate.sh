#!/bin/bash
# GPU box: insertion rate of the RNG-exact batched Widom replay on the reference's Henrys_coefficient deck (host driver), for a
# number of production cycles and of GPUs; prints the rate, <W> of the last block and the Henry coefficient line.
# usage: scripts/henry_rate.sh CYCLES [GPUS ...]
cd "$(dirname "$0")/.."
CYC=${1:-500000}; shift
for G in "${@:-1}"; do
  D=$(mktemp -d /tmp/henry.XXXX); cp -r oracle/_ref/examples/Henrys_coefficient/* $D/; chmod -R u+w $D
  out=$(timeout 600 ./graspa_b200/host/graspa_b200_mc $D --init 0 --equil 0 --prod $CYC --gpus $G $HENRY_FLAGS 2>&1)
  echo "$out" | grep '"cycles_per_s"' | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('gpus %d $HENRY_FLAGS: %.0f insertions/s (%d insertions in %.3f s, widom path %s, launches %d)' % (j.get('gpus',1), j['moves_per_s'], j['moves'], j['seconds'], j['widom_path'], j['kernel_launches']))"
  echo "$out" | grep -E "batched Widom replay" | sed 's/^/    /'
  echo "$out" | grep -E "Averaged Rosenbluth Weight" | tail -2 | sed 's/^/    /'
  echo "$out" | grep -iE "error|fail" | head -3
  rm -rf $D
done
