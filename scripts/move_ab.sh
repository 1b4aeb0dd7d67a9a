#!/bin/bash
# GPU box: A/B of the GCMC move path on the example decks: cycles/s of the host driver per setting.
# usage: scripts/move_ab.sh "ENV=.. ARGS" ...   (each argument: environment assignments, then driver flags starting with --)
cd "$(dirname "$0")/.."
run() {
  for deck in "XeKr-Mixture 20000" "CO2-MFI 3000" "CO2_NaX_Zeolite 5000"; do
    set -- $deck
    D=$(mktemp -d /tmp/mvab.XXXX); cp -r oracle/_ref/examples/$1/* $D/; chmod -R u+w $D
    out=$(env $VENV timeout 300 ./graspa_b200/host/graspa_b200_mc $D --init $2 --equil 0 --prod 0 $VARGS 2>&1)
    r=$(echo "$out" | grep '"cycles_per_s"' | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('%.0f cycles/s, %d launches, server starts %d moves %d' % (j['cycles_per_s'], j['kernel_launches'], j.get('server_starts', 0), j.get('server_moves', 0)))")
    e=$(echo "$out" | grep -E "^Total Energy|total energy|final_total|Final" | head -2 | tr '\n' ' ')
    echo "  [$VENV $VARGS] $1: $r"
    echo "$out" | grep -E "host time inside" | sed 's/^/      /'
    echo "$out" | grep -iE "error|fail|trap" | head -3
    rm -rf $D
  done
}
for V in "$@"; do
  VENV=""; VARGS=""
  for w in $V; do case "$w" in --*) VARGS="$VARGS $w";; *) VENV="$VENV $w";; esac; done
  run
done
