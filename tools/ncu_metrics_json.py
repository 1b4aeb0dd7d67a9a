#!/usr/bin/env python
"""From the raw page of an ncu capture of the Widom energy kernel (`ncu -i X.ncu-rep --page raw --csv`) to the two numbers bench.py
cannot measure itself: DRAM bytes per insertion and the executed FP64-pipe share (time-weighted over the captured launches).
Usage: tools/ncu_metrics_json.py raw.csv INSERTIONS_PER_CAPTURED_LAUNCH_SET source-description > profiles/r2_pair_kernel_metrics.json
Also writes the kept columns of the raw page next to it (profiles/r2_pair_kernel_ncu.csv): the committed evidence."""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
n_ins = int(sys.argv[2]); source = sys.argv[3]
col = {k: hdr.index(k) for k in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                 "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                                 "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "Grid Size", "Block Size")}


def to_bytes(v, u):
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


def to_ms(v, u):
    return float(v.replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[u]


t = [to_ms(r[col["gpu__time_duration.sum"]], units[col["gpu__time_duration.sum"]]) for r in data]
rd = [to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]) for r in data]
wr = [to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]]) for r in data]
fp = [float(r[col["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]]) for r in data]
iss = [float(r[col["smsp__issue_active.avg.pct_of_peak_sustained_active"]]) for r in data]
out = {"kernel": data[0][col["Kernel Name"]].split("(")[0], "launches_captured": len(data), "insertions_covered": n_ins,
       "kernel_ms_under_ncu": t, "dram_bytes_read": rd, "dram_bytes_written": wr,
       "dram_bytes_per_insertion": (sum(rd) + sum(wr)) / n_ins,
       "fp64_pipe_pct": sum(a * b for a, b in zip(t, fp)) / sum(t), "fp64_pipe_pct_per_launch": fp,
       "issue_active_pct": sum(a * b for a, b in zip(t, iss)) / sum(t), "source": source}
with open("profiles/r2_pair_kernel_ncu.csv", "w", newline="") as f:
    w = csv.writer(f); keep = sorted(col.values())
    w.writerow([hdr[i] for i in keep]); w.writerow([units[i] for i in keep])
    for r in data:
        w.writerow([r[i] for i in keep])
print(json.dumps(out, indent=1))
