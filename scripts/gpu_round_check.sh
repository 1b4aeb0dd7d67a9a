#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench, ncu launch list of the bench command, one full ncu capture of the Widom
# pair kernel and of the move kernel.  Everything lands in gpurun_out/ (scratch); summaries are copied to profiles/ by hand.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -c 3000 gpurun_out/bench_r1.json; tail -3 gpurun_out/bench_r1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv \
  python bench.py --steps 2 --warmup 3 --batch 40000 --no-cpu-baseline --no-secondary > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_widom_pair -s 3 -c 1 -o gpurun_out/prof_pair_r1 -f \
  python bench.py --steps 1 --warmup 3 --batch 40000 --no-cpu-baseline --no-secondary > gpurun_out/ncu_pair.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_widom_ewald -s 3 -c 1 -o gpurun_out/prof_ewald_r1 -f \
  python bench.py --steps 1 --warmup 3 --batch 40000 --no-cpu-baseline --no-secondary > gpurun_out/ncu_ewald.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_move -s 2000 -c 4 -o gpurun_out/prof_move_r1 -f \
  graspa_b200/host/graspa_b200_mc oracle/_ref/examples/CO2-MFI --init 200 > gpurun_out/ncu_move.log 2>&1
ls -la gpurun_out | tail -12
