#!/bin/bash
# kernel experiments: run the quick Widom bench (and the Widom parity tests) with alternative builds from build_dbg/
set -u
for lib in "$@"; do
  echo "=== $lib"
  GRASPA_B200_LIB="$PWD/$lib" timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "widom" 2>&1 | tail -1
  GRASPA_B200_LIB="$PWD/$lib" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j = json.loads(l); k = j['kernels']
        print('value %.0f ins/s  e2e %.0f  pair %.2f ms  ewald %.2f ms' % (j['value'], j['e2e']['value'], k['k_widom_pair_ms'], k['k_widom_ewald_ms']))
"
done
