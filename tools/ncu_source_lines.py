#!/usr/bin/env python
"""Per-source-line hot spots of one kernel from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.
Usage: tools/ncu_source_lines.py src.csv [top N]   -> markdown table: file:line, warp instructions, share, stall samples."""
import csv
import os
import sys


def main():
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = list(csv.reader(open(sys.argv[1])))
    cur = None; hdr = None; lines = []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = os.path.basename(r[1]); continue
        if r[0] == "Line No":
            hdr = r; continue
        if r[0] in ("Function Name", "Kernel Name") or hdr is None or cur is None:
            continue
        if r[2] != "-":            # SASS row under a source line
            continue
        try:
            inst = int(r[hdr.index("Instructions Executed")]); samp = int(r[hdr.index("# Samples")])
        except ValueError:
            continue
        d = dict(zip(hdr[4:], r[4:]))
        lines.append((cur, int(r[0]), r[1].strip(), inst, samp, d))
    tot_i = sum(l[3] for l in lines) or 1; tot_s = sum(l[4] for l in lines) or 1
    print(f"total warp instructions {tot_i}, stall samples {tot_s}\n")
    print("| file:line | warp inst | % inst | samples | % samples | top stalls | source |\n|---|---|---|---|---|---|---|")
    for f, ln, src, inst, samp, d in sorted(lines, key=lambda l: -l[4])[:top]:
        st = sorted(((int(v), k) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit()), reverse=True)[:3]
        sts = ", ".join(f"{k[6:]} {v}" for v, k in st if v > 0)
        print(f"| {f}:{ln} | {inst} | {100.0 * inst / tot_i:.1f} | {samp} | {100.0 * samp / tot_s:.1f} | {sts} | `{src[:90]}` |")


if __name__ == "__main__":
    main()
