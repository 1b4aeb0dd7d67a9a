"""Generates the committed golden fixtures under tests/golden/ (run in the build container only).

    python -m tests.golden.make_golden

Inputs: the reference's example decks under /root/reference/Examples and the reference's own routines through
oracle/_ref/libgraspa_ref_host.so (built by oracle/build_ref.sh from /root/reference/src_clean).
Outputs (small .npz files, committed):
  config_<X>.npz   box / force field / system arrays of the SURVEY section 8 configs A, B, D, E
                   + seeded trial batches with the REFERENCE harness' per-trial energies and flags,
                   + the reference's Ewald_Total energies and structure factors,
                   + the reference's rigid exclusion constants and tail corrections,
                   + oracle Widom insertions (W, energy terms, stage) on seeded randoms,
  rng_seed0.npz    the first 4096 values of the reference's uniform stream for srand(0).
The GPU-side tests never read /root/reference: they read these files.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from graspa_b200.types import System, TrialAtoms, species_counts, pseudo_atom_counts, CBMC_INSERTION  # noqa: E402
from oracle import oracle as orc                                                                      # noqa: E402
from tests.support.raspa_inputs import load_deck                                                      # noqa: E402

EX = "/root/reference/Examples"


def add_adsorbates(deck, comp, nmol, seed):
    """grow nmol molecules of `comp` with the oracle's CBMC insertion (always accepting successful growths)"""
    box, ff, s = deck["box"], deck["ff"], deck["system"]
    rng = np.random.default_rng(seed)
    ms = int(s.molsize[comp])
    alloc = s.alloc.copy(); alloc[comp] = max(alloc[comp], (nmol + 2) * ms)
    off_old = s.offsets
    n = int(alloc.sum())
    pos = np.zeros((n, 3)); q = np.zeros(n); ty = np.zeros(n, dtype=np.int64); mo = np.zeros(n, dtype=np.int64)
    sc = np.ones(n); scc = np.ones(n)
    off_new = np.concatenate([[0], np.cumsum(alloc)])
    for c in range(s.ncomp):
        k = int(s.alloc[c])
        sl_o = slice(int(off_old[c]), int(off_old[c]) + k); sl_n = slice(int(off_new[c]), int(off_new[c]) + k)
        pos[sl_n] = s.pos[sl_o]; q[sl_n] = s.charge[sl_o]; ty[sl_n] = s.type[sl_o]; mo[sl_n] = s.molid[sl_o]
    natoms = s.natoms.copy()
    cur = System(s.nhost, natoms, s.molsize, pos, q, ty, mo, sc, scc, alloc=alloc)
    placed = 0
    while placed < nmol:
        ws = orc.WidomSetup(box, ff, cur, comp, deck["beta"], deck["ntrials"], deck["norient"],
                            np.zeros(2 * max(box.nvec, 1)), np.zeros(2 * max(box.nvec, 1)))
        out, stage, sel, mpos, _ = orc.widom_insertion(ws, rng.random((deck["ntrials"], 3)), rng.random(), rng.random((deck["norient"], 3)), rng.random())
        if stage != 0 or out[1] + out[3] > 0.0:
            continue
        o = int(cur.offsets[comp]); base = o + int(cur.natoms[comp])
        tmpl = slice(o, o + ms)
        cur.pos[base:base + ms] = mpos
        cur.charge[base:base + ms] = cur.charge[tmpl]; cur.type[base:base + ms] = cur.type[tmpl]
        cur.molid[base:base + ms] = int(cur.natoms[comp]) // ms
        cur.natoms[comp] += ms
        placed += 1
    deck["system"] = cur
    return deck


def save_config(name, deck, comp, seed, ntb=48, nwidom=24):
    box, ff, s = deck["box"], deck["ff"], deck["system"]
    rng = np.random.default_rng(seed)
    ms = int(s.molsize[comp]); cs = ms - 1
    new_molid = int(s.natoms[comp]) // ms
    out = dict(
        cell=box.cell, alpha=box.alpha, kmax=np.array(box.kmax), recip_cutoff=box.recip_cutoff, prefactor=box.prefactor,
        eps=ff.epsilon, sigma=ff.sigma, shift=ff.shift, cutoff_vdw=ff.cutoff_vdw, cutoff_coul=ff.cutoff_coul, overlap=ff.overlap,
        no_charges=int(ff.no_charges), use_tail=ff.use_tail, tail_energy=ff.tail_energy,
        nhost=s.nhost, natoms=s.natoms, molsize=s.molsize, alloc=s.alloc, pos=s.pos, charge=s.charge, type=s.type, molid=s.molid,
        beta=deck["beta"], temperature=deck["temperature"], ntrials=deck["ntrials"], norient=deck["norient"], comp=comp,
    )
    # ---- trial batch 1: first-bead style (chainsize 1)
    rnd = rng.random((ntb, 3))
    t1 = orc.trial_positions(box, s, CBMC_INSERTION, comp, 0, ntb, rnd)
    e1, f1, c1 = orc.ref_trial_energies(box, ff, s, ntb, 1, t1, comp, new_molid)
    out.update(tb1_pos=t1.pos, tb1_charge=t1.charge, tb1_type=t1.type, tb1_energy=e1, tb1_flag=f1, tb1_counts=c1)
    # ---- trial batch 2: chain style (chainsize ms-1) around a favourable first bead
    if cs > 0:
        best = int(np.argmin(np.where(f1 == 0, e1.sum(axis=1), np.inf)))
        rnd2 = rng.random((ntb, 3))
        t2 = orc.trial_orientations(s, CBMC_INSERTION, comp, 1, cs, ntb, rnd2, t1.pos[best])
        e2, f2, c2 = orc.ref_trial_energies(box, ff, s, ntb, cs, t2, comp, new_molid)
        out.update(tb2_pos=t2.pos, tb2_charge=t2.charge, tb2_type=t2.type, tb2_energy=e2, tb2_flag=f2, tb2_counts=c2, tb2_cs=cs)
    # ---- Ewald total + structure factors + exclusion constants from the reference
    if not ff.no_charges:
        E, sa, sf = orc.ref_ewald_total(box, s)
        o = int(s.offsets[comp])
        ex = orc.ref_exclusion_rigid(box, s.pos[o:o + ms], s.charge[o:o + ms], s.scale_coul[o:o + ms])
        out.update(ewald_E=E, sf_ads=sa, sf_fw=sf, excl=np.array(ex))
    # ---- tail from the reference
    npseudo = pseudo_atom_counts(s, ff.ntypes)
    counts = [species_counts(s, c, ff.ntypes) for c in range(s.ncomp)]
    out.update(npseudo=npseudo, tail_total=orc.ref_tail_total(ff, npseudo, box.volume),
               tail_ins=np.array([orc.ref_tail_difference(ff, npseudo, box.volume, counts, c, True) for c in range(s.nhost, s.ncomp)]),
               tail_del=np.array([orc.ref_tail_difference(ff, npseudo, box.volume, counts, c, False) for c in range(s.nhost, s.ncomp)]))
    if s.ncomp - s.nhost >= 2:
        out.update(tail_swap=orc.ref_tail_identity_swap(ff, npseudo, box.volume, counts, s.nhost, s.nhost + 1))
    # ---- Widom insertions through the oracle (itself pinned against the reference routines by tests/test_oracle_vs_ref.py)
    ws = orc.WidomSetup(box, ff, s, comp, deck["beta"], deck["ntrials"], deck["norient"],
                        out.get("sf_ads"), out.get("sf_fw"))
    wr = rng.random((nwidom, deck["ntrials"] + deck["norient"], 3)); wu = rng.random((nwidom, 2))
    w8, wst, wc = orc.widom_batch(ws, wr, wu)
    out.update(widom_rnd=wr, widom_uni=wu, widom_out=w8, widom_stage=wst, widom_counts=wc)
    np.savez_compressed(os.path.join(HERE, f"config_{name}.npz"), **out)
    print(f"config_{name}: N={int(s.natoms.sum())} nvec={box.nvec} <W>={w8[:, 0].mean():.6g} fails={(wst > 0).sum()} "
          f"size={os.path.getsize(os.path.join(HERE, f'config_{name}.npz')) / 1024:.0f} KB")


def main():
    # A: Henrys_coefficient as shipped (Mg-MOF-74 5x3x3, CO2 Widom)
    save_config("A", load_deck(f"{EX}/Henrys_coefficient"), 1, 1234)
    # E: the synthetic 4x4x4 supercell of the same deck (BASELINE.json config 5)
    save_config("E", load_deck(f"{EX}/Henrys_coefficient", unitcells=(4, 4, 4)), 1, 1235)
    # B: CO2-MFI with 20 CO2 molecules grown by the oracle (exercises guest-guest terms and exclusions)
    save_config("B", add_adsorbates(load_deck(f"{EX}/CO2-MFI", extra_alloc=0), 1, 20, 77), 1, 1236)
    # D: Xe/Kr mixture: two monatomic adsorbates, no charges, O-O tail correction; 12 Kr + 12 Xe
    d = load_deck(f"{EX}/XeKr-Mixture")
    d = add_adsorbates(d, 1, 12, 78); d = add_adsorbates(d, 2, 12, 79)
    save_config("D", d, 2, 1237)
    np.savez_compressed(os.path.join(HERE, "rng_seed0.npz"), u=orc.ref_uniform_stream(0, 4096))


if __name__ == "__main__":
    main()
