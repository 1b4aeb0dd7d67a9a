#!/bin/bash
# One GPU-box pass (round 2): parity tests, smoke, bench, ncu launch list of the bench command, full ncu captures of the Widom
# energy kernel (cell-sorted stage), the Widom Fourier kernel and the move kernel.  Everything lands in gpurun_out/ (scratch);
# summaries are copied to profiles/ by hand (tools/ncu_summary.py).
set -u
R=${1:-r2}
WHAT=${2:-all}     # all | check | prof
mkdir -p gpurun_out
if [ "$WHAT" != prof ]; then
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/${R}_gputest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
# the committed N = 1 sums of the 10^7 job are rewritten with the kernels of this build (one pass), then the bench checks itself against them
timeout 600 python bench.py --write-job-sums --no-secondary --no-cpu-baseline --steps 1 --warmup 1 > /dev/null 2>&1; cp tests/golden/job_sums_E.json gpurun_out/job_sums_E.json
timeout 900 python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; tail -c 1500 gpurun_out/${R}_bench.json; tail -3 gpurun_out/${R}_bench.err
for deck in "XeKr-Mixture 20000" "CO2-MFI 3000"; do
  set -- $deck
  D=$(mktemp -d /tmp/rc.XXXX); cp -r oracle/_ref/examples/$1/* $D/; chmod -R u+w $D
  ./graspa_b200/host/graspa_b200_mc $D --init $2 --equil 0 --prod 0 2>&1 | grep -E "cycles_per_s|host time" | tee -a gpurun_out/${R}_moves.log
  ./graspa_b200/host/graspa_b200_mc $D --init $2 --equil 0 --prod 0 --no-server 2>&1 | grep -E "cycles_per_s|host time" | sed 's/^/[--no-server] /' | tee -a gpurun_out/${R}_moves.log
  ./graspa_b200/host/graspa_b200_mc $D --init $2 --equil 0 --prod 0 --timing 2>&1 | grep -E "device time" | tee -a gpurun_out/${R}_moves.log
  rm -rf $D
done
fi
[ "$WHAT" = check ] && exit 0
# the bench's own command (the 10^7-insertion job in sub-batches of 10^6), without the CPU legs and the GCMC section
NCUB="python bench.py --no-cpu-baseline --no-secondary"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${R}_launches.csv \
  $NCUB --steps 1 --warmup 3 > gpurun_out/${R}_bench_under_ncu.log 2>&1
# the two energy launches of one 10^6-insertion sub-batch of the timed step (20 launches per pass over the job: skip the 3 warm-up passes)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_wc_energy_lt -s 60 -c 2 -o gpurun_out/prof_wc_energy_${R} -f \
  $NCUB --steps 1 --warmup 3 > gpurun_out/${R}_ncu_wc.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_widom_ewald -s 30 -c 1 -o gpurun_out/prof_ewald_${R} -f \
  $NCUB --steps 1 --warmup 3 > gpurun_out/${R}_ncu_ewald.log 2>&1
D=$(mktemp -d /tmp/rc.XXXX); cp -r oracle/_ref/examples/XeKr-Mixture/* $D/; chmod -R u+w $D
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_move -s 4000 -c 3 -o gpurun_out/prof_move_${R} -f \
  graspa_b200/host/graspa_b200_mc $D --init 6000 --equil 0 --prod 0 --no-server > gpurun_out/${R}_ncu_move.log 2>&1   # per-launch k_move: a resident kernel that waits for host commands cannot be replayed by ncu
ls -la gpurun_out | tail -12
