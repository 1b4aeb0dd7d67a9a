#!/bin/bash
# Refresh of the committed evidence after the short erfc table went into the Widom energy kernel (round 2, final): job sums of this build,
# Widom parity tests, full ncu capture of the two energy launches of one sub-batch, launch list of the bench command, bench line.
set -u
mkdir -p gpurun_out
timeout 600 python bench.py --write-job-sums --no-secondary --no-cpu-baseline --steps 1 --warmup 1 > /dev/null 2>&1; cp tests/golden/job_sums_E.json gpurun_out/job_sums_E.json
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_config_c.py -m gpu -q -x 2>&1 | tail -2
NCUB="python bench.py --no-cpu-baseline --no-secondary"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wc_energy_lt -s 60 -c 2 -o gpurun_out/prof_wc_energy_r2f -f $NCUB --steps 1 --warmup 3 > gpurun_out/r2f_ncu_wc.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2f_launches.csv $NCUB --steps 1 --warmup 3 > gpurun_out/r2f_bench_under_ncu.log 2>&1
timeout 900 python bench.py > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; tail -c 300 gpurun_out/r2f_bench_n1.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2f_bench_n1.json").read().strip().splitlines()[-1])
print({k:j[k] for k in ("value","ms_per_step","gpu_launches","job_check")}); print(j["e2e"]["value"], j["roofline"]["frac"], j["kernels"])
PY
ls -la gpurun_out | grep r2f
