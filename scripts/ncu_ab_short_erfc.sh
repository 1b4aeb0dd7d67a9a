M=gpu__time_duration.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,launch__grid_size,launch__block_size,smsp__sass_inst_executed_op_shared_ld.sum,smsp__thread_inst_executed_per_inst_executed.ratio
for v in 0 1; do
  if [ $v = 1 ]; then export GB_WC_NO_SHORT_ERFC=1; fi
  ncu --metrics $M --clock-control none -k regex:k_wc_energy_lt -s 60 -c 2 --csv --log-file gpurun_out/ab_short_$v.csv python bench.py --no-cpu-baseline --no-secondary --steps 1 --warmup 3 > /dev/null 2>&1
done
python - <<'PY'
import csv
for v in (0,1):
    rows=[r for r in csv.reader(open(f"gpurun_out/ab_short_{v}.csv")) if len(r)>10]
    hdr=rows[0]; 
    for r in rows[1:]:
        d=dict(zip(hdr,r))
        print(v, d.get("ID"), d.get("Kernel Name","")[:40], d.get("Metric Name"), d.get("Metric Value"))
PY
