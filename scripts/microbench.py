#!/usr/bin/env python
"""Kernel microbenchmarks of SURVEY.md section 8(d), items 1 and 4, on one B200 through the C ABI:

  1. pair kernel: framework of config A (2430 atoms), B (2304 + CO2), C (NaX + movable Na+), D (1216, no charges), E (3456); trial batches of
     T in {10, 10^3, 10^5, 10^6} groups of 1 atom or of a 3-atom CO2; first-bead positions uniform in the cell, orientations from
     uniform randoms; seed 1234.  Device time by CUDA events on the engine's stream (gb_timing_read), algorithmic flops by
     SURVEY's count F = N_pairs (44 | 20) + 17 N_vdw + 9 N_coul + 2 N_in with the in-cutoff fractions the CPU oracle counts
     on a sample of the same batch.
  4. Fourier kernels: the Widom Fourier kernel per insertion and the single-move delta per call at the nvec of each charged
     config, 3 moved atoms (insertion) and 6 (translation: 3 old + 3 new).

    python scripts/microbench.py > profiles/r2_microbench.json        (oracle/ is used as the counter of in-cutoff pairs only)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graspa_b200 import engine  # noqa: E402
from graspa_b200.types import TrialAtoms  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from tests.conftest import load_config  # noqa: E402


def rotations(rng, n):
    q = rng.normal(size=(n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    a, b, c, d = q.T
    return np.stack([np.stack([a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)], -1),
                     np.stack([2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b)], -1),
                     np.stack([2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d], -1)], 1)


def main():
    orc.build()
    rng = np.random.default_rng(1234)
    out = {"what": "SURVEY section 8(d) microbenchmarks", "pair": [], "fourier": []}
    peak = None
    for name in ("A", "B", "C", "D", "E"):
        box, ff, s, z = load_config(name)
        comp = int(z["comp"]); ms = int(s.molsize[comp]); o = int(s.offsets[comp])
        eng = engine.Engine(0).setup(box, ff, s, float(z["beta"]), 10, 10)
        if peak is None:
            peak = eng.measure_fp64_peak()
            out["fp64_peak_tflops"] = peak
        charged = not ff.no_charges
        if "sf_ads" in z and charged:
            eng.upload_structure_factors(z["sf_ads"], z["sf_fw"])
        nsys = int(sum(int(s.natoms[c]) for c in range(s.ncomp)))
        tmpl_pos = s.pos[o:o + ms] - s.pos[o]; tmpl_q = s.charge[o:o + ms]; tmpl_t = s.type[o:o + ms]
        cell = box.cell.reshape(3, 3)
        per_pair = 20.0 if box.cubic else 44.0
        eng.timing_enable(True)
        for cs in sorted({1, ms}):
            for T in (10, 1000, 100000, 1000000):
                first = rng.random((T, 3)) @ cell
                if cs == 1:
                    pos = first; q = np.full(T, tmpl_q[0]); ty = np.full(T, tmpl_t[0])
                else:
                    R = rotations(rng, T)
                    pos = (first[:, None, :] + np.einsum("nij,aj->nai", R, tmpl_pos)).reshape(-1, 3)
                    q = np.tile(tmpl_q, T); ty = np.tile(tmpl_t, T)
                tr = TrialAtoms(pos, q, ty)
                # in-cutoff fractions from the oracle on a sample of the batch
                ns = min(T, 1500)
                trs = TrialAtoms(pos[:ns * cs], q[:ns * cs], ty[:ns * cs])
                _, _, cnt = orc.trial_energies(box, ff, s, ns, cs, trs, comp, 10 ** 9)
                npairs_s = float(cnt[0]); f_vdw = cnt[1] / npairs_s; f_coul = cnt[2] / npairs_s; f_in = cnt[3] / npairs_s
                reps = 5 if T <= 100000 else 2
                eng.trial_energies(T, cs, tr, comp, 10 ** 9)
                l0 = eng.launch_count()
                eng.trial_energies(T, cs, tr, comp, 10 ** 9)
                route = "cell-sorted (k_wc_energy_lt)" if eng.launch_count() - l0 > 1 else "k_trial_energies"
                eng.timing_read(0, reset=True)
                t0 = time.perf_counter()
                for _ in range(reps):
                    eng.trial_energies(T, cs, tr, comp, 10 ** 9)
                wall = (time.perf_counter() - t0) / reps
                ms_dev, nl = eng.timing_read(0, reset=True)
                ms_dev /= reps
                npairs = float(T) * cs * nsys
                flops = npairs * (per_pair + 17.0 * f_vdw + 9.0 * f_coul + 2.0 * f_in)
                out["pair"].append({"config": name, "system_atoms": nsys, "cell": "orthorhombic" if box.cubic else "triclinic", "charged": charged,
                                    "kernel": "gb_trial_energies: " + route, "trial_groups": T, "atoms_per_group": cs,
                                    "device_ms": ms_dev, "call_ms_with_copies": wall * 1e3, "pairs_per_s": npairs / (ms_dev * 1e-3),
                                    "in_cutoff_fraction": {"vdw": f_vdw, "coulomb": f_coul}, "algorithmic_tflops": flops / (ms_dev * 1e-3) / 1e12,
                                    "frac_of_fp64_peak": flops / (ms_dev * 1e-3) / 1e12 / peak})
        # batched Widom (tile-culled pair kernel + Fourier kernel) where the component is a 3-atom molecule
        if ms > 1:
            ws = orc.WidomSetup(box, ff, s, comp, float(z["beta"]), 10, 10, z["sf_ads"], z["sf_fw"])
            eng.set_exclusion_constants(comp, *ws.excl)
            for n in (1000, 100000):
                rnd = rng.random((n * 20, 3)); uni = rng.random((n, 2))
                eng.widom_batch(comp, rnd, uni, want_outputs=False)
                eng.timing_read(2, reset=True)                      # a reset clears both families
                eng.widom_batch(comp, rnd, uni, want_outputs=False)
                mp, _ = eng.timing_read(0); me, _ = eng.timing_read(1, reset=True)
                out["pair"].append({"config": name, "kernel": "k_widom_pair (gb_widom_batch)", "insertions": n, "device_ms": mp,
                                    "insertions_per_s_pair_kernel": n / (mp * 1e-3)})
                if charged:
                    out["fourier"].append({"config": name, "nvec": int(box.nvec), "kernel": "k_widom_ewald (gb_widom_batch)", "moved_atoms": ms,
                                           "insertions": n, "device_ms": me, "ns_per_insertion": me * 1e6 / n})
        if charged and ms > 1:
            # single-move Fourier delta: 3 new atoms (insertion) and 3 old + 3 new (translation)
            p_new = (rng.random(3) @ cell) + tmpl_pos
            for nold, nnew in ((0, ms), (ms, ms)):
                pos = np.concatenate([s.pos[o:o + ms]] * (1 if nold else 0) + [p_new]) if nold else p_new
                qq = np.tile(tmpl_q, 2 if nold else 1)
                eng.ewald_delta_explicit(False, nold, nnew, pos, qq, np.ones(len(qq)))
                eng.timing_read(1, reset=True)
                reps = 200
                t0 = time.perf_counter()
                for _ in range(reps):
                    eng.ewald_delta_explicit(False, nold, nnew, pos, qq, np.ones(len(qq)))
                wall = (time.perf_counter() - t0) / reps
                me, nl = eng.timing_read(1, reset=True)
                out["fourier"].append({"config": name, "nvec": int(box.nvec), "kernel": "k_ewald_delta (gb_ewald_delta_explicit)", "moved_atoms": nold + nnew,
                                       "device_us_per_call": me * 1e3 / reps, "call_us_with_copies_and_sync": wall * 1e6})
        eng.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
