#!/usr/bin/env python
"""Per-kernel SASS instruction counts of libgraspa_b200.so (cuobjdump -sass): total, FP64 (DFMA/DMUL/DADD/DSETP/MUFU.*64),
shared-memory loads/stores, TMA bulk copies (UBLKCP), mbarrier operations (SYNCS), cluster / DSMEM instructions.
Usage: tools/sass_counts.py [lib]  -> markdown table"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "graspa_b200/libgraspa_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
fn = None
cnt = collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = m.group(1); cnt[fn] = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and fn:
        op = m.group(1); c = cnt[fn]
        c["total"] += 1
        base = op.split(".")[0]
        if base in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX") or (base == "MUFU" and "64" in op): c["fp64"] += 1
        if base == "DFMA": c["dfma"] += 1
        if base in ("LDS", "LDSM"): c["lds"] += 1
        if base == "STS": c["sts"] += 1
        if base == "UBLKCP": c["ublkcp"] += 1
        if base == "SYNCS": c["syncs"] += 1
        if base in ("UTMALDG", "UTMASTG"): c["utma"] += 1
        if base in ("UCGABAR_ARV", "UCGABAR_WAIT", "CGAERRBAR") or "CGA" in op: c["cluster"] += 1
        if base in ("LDG", "LD") and ("128" in op or "64" in op): c["ldg_wide"] += 1
        if base == "SHFL": c["shfl"] += 1
print(f"`{lib}`: architectures {', '.join(arch)}\n")
print("| kernel | SASS instr | FP64 | DFMA | LDS | STS | SHFL | wide LDG | UBLKCP | SYNCS | cluster |\n|---|---|---|---|---|---|---|---|---|---|---|")
import subprocess as sp
for fn, c in cnt.items():
    try:
        name = sp.run(["c++filt", fn], capture_output=True, text=True).stdout.strip(); name = re.sub(r"\(anonymous namespace\)::", "", name).split("(")[0]
    except Exception:
        name = fn
    print(f"| {name} | {c['total']} | {c['fp64']} | {c['dfma']} | {c['lds']} | {c['sts']} | {c['shfl']} | {c['ldg_wide']} | {c['ublkcp']} | {c['syncs']} | {c['cluster']} |")
