#!/usr/bin/env python
"""profiles/r2_microbench.json (scripts/microbench.py) -> the markdown tables of profiles/r2_microbench.md"""
import json
import sys

d = json.load(open(sys.argv[1] if len(sys.argv) > 1 else "profiles/r2_microbench.json"))
print("# Kernel microbenchmarks of SURVEY section 8(d), round 2 (scripts/microbench.py, one B200)\n")
print(f"FP64 DFMA peak measured in the same process: {d['fp64_peak_tflops']:.1f} TFLOP/s.  `frac` = ALGORITHMIC flops (every system-atom x trial-atom pair at 44 flops, "
      "20 in an orthorhombic cell, plus the in-cutoff terms) / device time / that peak; the cell-sorted route skips most minimum-image tests, so its algorithmic figure "
      "can exceed 1.\n")
print("## gb_trial_energies: caller-supplied trial atoms (groups of 1 atom or of the 3-atom CO2; config D: Xe)\n")
print("Round 1 evaluated every batch with one CTA per group against all system atoms (k_trial_energies: 0.32 of the peak on E at 10^5-10^6 groups, 0.14 on B / D).  "
      "Round 2 sends batches of >= 64 trial atoms per 2 A cell (equal groups, nothing of the system excluded) to the cell-sorted energy kernel of the Widom stage.\n")
print("| config | system atoms | atoms per group | groups | route | device ms | G pairs/s | algorithmic TFLOP/s | frac of FP64 peak |\n|---|---|---|---|---|---|---|---|---|")
for r in d["pair"]:
    if "trial_groups" in r:
        print(f"| {r['config']} | {r['system_atoms']} | {r['atoms_per_group']} | {r['trial_groups']} | {r['kernel'].split(': ')[1]} | {r['device_ms']:.3f} | "
              f"{r['pairs_per_s'] / 1e9:.1f} | {r['algorithmic_tflops']:.2f} | {r['frac_of_fp64_peak']:.2f} |")
print("\n## gb_widom_batch pair stage alone\n\n| config | insertions | device ms | M insertions/s (pair stage) |\n|---|---|---|---|")
for r in d["pair"]:
    if "insertions" in r:
        print(f"| {r['config']} | {r['insertions']} | {r['device_ms']:.3f} | {r['insertions_per_s_pair_kernel'] / 1e6:.2f} |")
print("\n## Fourier kernels\n\n| config | nvec | kernel | moved atoms | figure |\n|---|---|---|---|---|")
for r in d["fourier"]:
    fig = (f"{r['ns_per_insertion']:.1f} ns per insertion ({r['insertions']} insertions)" if "ns_per_insertion" in r
           else f"{r['device_us_per_call']:.1f} us per call on the device, {r['call_us_with_copies_and_sync']:.1f} us with copies and sync")
    print(f"| {r['config']} | {r['nvec']} | {r['kernel']} | {r['moved_atoms']} | {fig} |")
