#!/bin/bash
# GPU box: A/B of the Widom pair stages and of the cell-sorted stage's knobs on the bench workload (config E, 400 000 insertions).
# Each line: the knobs, then value (insertions/s, inputs resident), ms per step and the pair / Fourier stage times.
cd "$(dirname "$0")/.."
run() {
  echo -n "$* : "
  env "$@" python bench.py --no-cpu-baseline --no-secondary --steps 3 --warmup 3 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%.3f M ins/s, %.2f ms/step, e2e %.3f M, pair %.2f ms, fourier %.2f ms, launches %d' % (j['value']/1e6, j['ms_per_step'], j['e2e']['value']/1e6, j['kernels'].get('pair_stage_ms', j['kernels'].get('k_widom_pair_ms', 0)), j['kernels']['k_widom_ewald_ms'], j['gpu_launches']))"
}
for v in "$@"; do run $v; done
