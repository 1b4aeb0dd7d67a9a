#!/usr/bin/env python
"""Summarise an ncu report (raw page exported with `ncu -i X.ncu-rep --page raw --csv`) into the short metric table kept
under profiles/.  Usage: tools/ncu_summary.py raw.csv [launch index ...]"""
import csv
import re
import sys

KEEP = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smsp__sass_inst_executed_op_shared_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
]
STALL = re.compile(r"smsp__average_warps?_issue_stalled_(\w+)_per_issue_active\.ratio|smsp__average_warp_latency_issue_stalled_(\w+)\.ratio")


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    pick = [int(a) for a in sys.argv[2:]] or list(range(len(data)))
    for n in pick:
        row = data[n]
        print(f"### launch {n}\n")
        print("| metric | unit | value |\n|---|---|---|")
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                print(f"| {k} | {units[i]} | {row[i]} |")
        stalls = []
        for i, k in enumerate(hdr):
            m = STALL.match(k)
            if m:
                try:
                    stalls.append((float(row[i]), k))
                except ValueError:
                    pass
        for v, k in sorted(stalls, reverse=True)[:8]:
            print(f"| {k} | ratio | {v:.3f} |")
        print()


if __name__ == "__main__":
    main()
