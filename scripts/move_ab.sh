#!/bin/bash
# GPU box: A/B of the one-kernel move path (k_move) on the GCMC example decks: cycles/s of the host driver per knob setting.
cd "$(dirname "$0")/.."
run() {
  for deck in "XeKr-Mixture 20000" "CO2-MFI 3000" "CO2_NaX_Zeolite 5000"; do
    set -- $deck
    D=$(mktemp -d /tmp/mvab.XXXX); cp -r oracle/_ref/examples/$1/* $D/; chmod -R u+w $D
    r=$(env $VARIANT ./graspa_b200/host/graspa_b200_mc $D --init $2 --equil 0 --prod 0 2>&1 | grep '"cycles_per_s"' | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('%.0f cycles/s' % j['cycles_per_s'])")
    echo "  [$VARIANT] $1: $r"
    rm -rf $D
  done
}
for VARIANT in "$@"; do run; done
