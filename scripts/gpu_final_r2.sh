#!/bin/bash
# Final GPU pass of round 2: the whole GPU suite, smoke, ncu capture + launch list of the bench command with the final kernels, bench line.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r2g_gputest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
NCUB="python bench.py --no-cpu-baseline --no-secondary"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wc_energy_lt -s 60 -c 2 -o gpurun_out/prof_wc_energy_r2g -f $NCUB --steps 1 --warmup 3 > gpurun_out/r2g_ncu_wc.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2g_launches.csv $NCUB --steps 1 --warmup 3 > gpurun_out/r2g_bench_under_ncu.log 2>&1
timeout 900 python bench.py > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err; tail -c 300 gpurun_out/r2g_bench_n1.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2g_bench_n1.json").read().strip().splitlines()[-1])
print({k:j[k] for k in ("value","ms_per_step","gpu_launches","job_check")}); print(j["e2e"]["value"], j["roofline"]["frac"], j["kernels"])
print({k:(v["ours"],v["ratio"],v["results_match"]) for k,v in j["vs_reference_cuda"].items()}); print(j["multibox_xekr"]["value"])
PY
