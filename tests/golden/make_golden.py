"""Generates the committed golden fixtures under tests/golden/ (run in the build container only).

    python -m tests.golden.make_golden

Inputs: the reference's example decks under /root/reference/Examples and the reference's own routines through
oracle/_ref/libgraspa_ref_host.so (built by oracle/build_ref.sh from /root/reference/src_clean).
Outputs (small .npz files, committed):
  config_<X>.npz   box / force field / system arrays of the SURVEY section 8 configs A, B, C, D, E
                   + seeded trial batches with the REFERENCE harness' per-trial energies and flags,
                   + the reference's Ewald_Total energies and structure factors,
                   + the reference's rigid exclusion constants and tail corrections,
                   + oracle Widom insertions (W, energy terms, stage) on seeded randoms,
                   + seeded translation / rotation moves of every movable component (separated framework components
                     included) with the REFERENCE's single-body deltas (new - old through the harness' pair loop) and
                     Fourier deltas (Ewald_Total after - Ewald_Total before), and one insertion likewise,
                   + (config C) the block-pocket list,
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


def load_deck_via_host(folder):
    """Decks with separated framework components, mixing-rule overrides and block pockets (config C) are parsed by the C++ host
    reader (graspa_b200_mc --dump-deck, no GPU needed), whose parse of this deck is checked move by move against the reference
    program (tests/test_gpu_trace_parity.py); the arrays are inputs, every energy stored below comes from the reference harness."""
    import json, subprocess
    from graspa_b200.types import Box, ForceField
    exe = os.path.join(HERE, "..", "..", "graspa_b200", "host", "graspa_b200_mc")
    d = json.loads(subprocess.check_output([exe, "--dump-deck", folder]))
    box = Box(np.array(d["cell"]), alpha=d["alpha"], kmax=tuple(d["kmax"]), recip_cutoff=d["recip_cutoff"], prefactor=d["prefactor"])
    ff = ForceField(d["eps"], d["sigma"], d["shift"], d["cutoff_vdw"], d["cutoff_coul"], overlap=d["overlap"], no_charges=bool(d["no_charges"]),
                    use_tail=np.array(d["use_tail"]), tail_energy=np.array(d["tail_energy"]))
    natoms, molsize, alloc, P, Q, T, M = [], [], [], [], [], [], []
    for f in d["framework"]:
        n = len(f["type"])
        natoms.append(n); molsize.append(int(f["molsize"])); alloc.append(n)
        P.append(np.array(f["pos"]).reshape(n, 3)); Q.append(np.array(f["charge"])); T.append(np.array(f["type"], dtype=np.int64))
        M.append(np.array(f.get("molid", [0] * n), dtype=np.int64))
    for a in d["adsorbates"]:
        ms = len(a["type"])
        natoms.append(0); molsize.append(ms); alloc.append(ms)
        P.append(np.array(a["pos"]).reshape(ms, 3)); Q.append(np.array(a["charge"])); T.append(np.array(a["type"], dtype=np.int64)); M.append(np.zeros(ms, dtype=np.int64))
    system = System(len(d["framework"]), np.array(natoms), np.array(molsize), np.concatenate(P), np.concatenate(Q), np.concatenate(T),
                    np.concatenate(M), alloc=np.array(alloc))
    pockets = [(np.array(a["pocket_centers"]).reshape(-1, 3), np.array(a["pocket_radii"]), int(a["invert_pockets"])) for a in d["adsorbates"]]
    return dict(box=box, ff=ff, system=system, beta=d["beta"], temperature=d["temperature"], names=d["names"],
                ntrials=d["ntrials"], norient=d["norient"], pockets=pockets)


def _random_rotation(rng):
    q = rng.normal(size=4); q /= np.linalg.norm(q)
    a, b, c, d = q
    return np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                     [2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b)],
                     [2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d]])


def _with_molecule(s, comp, mol, newpos):
    """copy of the system with molecule `mol` of `comp` at newpos (mol == number of molecules: appended)"""
    ms = int(s.molsize[comp]); o = int(s.offsets[comp])
    t = System(s.nhost, s.natoms.copy(), s.molsize.copy(), s.pos.copy(), s.charge.copy(), s.type.copy(), s.molid.copy(),
               s.scale.copy(), s.scale_coul.copy(), alloc=s.alloc.copy())
    if mol * ms >= int(s.natoms[comp]):
        assert (mol + 1) * ms <= int(s.alloc[comp])
        t.charge[o + mol * ms:o + (mol + 1) * ms] = s.charge[o:o + ms]; t.type[o + mol * ms:o + (mol + 1) * ms] = s.type[o:o + ms]
        t.molid[o + mol * ms:o + (mol + 1) * ms] = mol
        t.natoms[comp] += ms
    t.pos[o + mol * ms:o + (mol + 1) * ms] = newpos
    return t


def reference_move_deltas(box, ff, s, rng, nper=4):
    """Seeded translation+rotation moves of every component that has movable molecules (molsize < natoms, i.e. not the rigid
    framework block), each with the REFERENCE's answers:
      sb_delta  [HHv, HHr, HGv, HGr, GGv, GGr] = ref_trial_energies(new) - ref_trial_energies(old)   (pair body VDW_Coulomb.cu:741-818)
      sb_ewald  [same-type, cross] = Ewald_Total(after) - Ewald_Total(before)   (ewald_preparation.h:5-259; GG includes HH, :174)
      sb_temp   structure factors of the moved species after the move"""
    rows = dict(comp=[], mol=[], old=[], new=[], delta=[], flag=[], ewald=[], temp=[])
    E0 = sa0 = sf0 = None
    if not ff.no_charges:
        E0, sa0, sf0 = orc.ref_ewald_total(box, s)
    for c in range(s.ncomp):
        ms = int(s.molsize[c]); nm = int(s.natoms[c]) // max(ms, 1)
        if nm < 2 and c < s.nhost:
            continue
        if nm == 0:
            continue
        o = int(s.offsets[c])
        for _ in range(nper):
            mol = int(rng.integers(0, nm))
            sl = slice(o + mol * ms, o + (mol + 1) * ms)
            old = s.pos[sl].copy()
            R = _random_rotation(rng) if ms > 1 else np.eye(3)
            new = old[0] + (old - old[0]) @ R.T + (rng.random(3) - 0.5) * 1.2
            ta_old = TrialAtoms(old, s.charge[sl], s.type[sl]); ta_new = TrialAtoms(new, s.charge[sl], s.type[sl])
            eo, fo, _ = orc.ref_trial_energies(box, ff, s, 1, ms, ta_old, c, mol)
            en, fn, _ = orc.ref_trial_energies(box, ff, s, 1, ms, ta_new, c, mol)
            d4 = en[0] - eo[0]                       # harness columns: {host vdw, host real, adsorbate vdw, adsorbate real}
            d6 = np.array([d4[0], d4[1], d4[2], d4[3], 0.0, 0.0]) if c < s.nhost else np.array([0.0, 0.0, d4[0], d4[1], d4[2], d4[3]])
            ew = np.zeros(2); temp = np.zeros(0)
            if not ff.no_charges:
                E1, sa1, sf1 = orc.ref_ewald_total(box, _with_molecule(s, c, mol, new))
                if c < s.nhost:
                    ew = np.array([E1[1] - E0[1], E1[2] - E0[2]]); temp = sf1
                else:
                    ew = np.array([E1[0] - E0[0], E1[2] - E0[2]]); temp = sa1
            rows["comp"].append(c); rows["mol"].append(mol); rows["old"].append(old.ravel()); rows["new"].append(new.ravel())
            rows["delta"].append(d6); rows["flag"].append(int(fn[0])); rows["ewald"].append(ew); rows["temp"].append(temp)
    if not rows["comp"]:
        return {}
    msmax = max(len(x) for x in rows["old"])
    pad = lambda L: np.array([np.concatenate([x, np.zeros(msmax - len(x))]) for x in L])
    return dict(sb_comp=np.array(rows["comp"]), sb_mol=np.array(rows["mol"]), sb_old=pad(rows["old"]), sb_new=pad(rows["new"]),
                sb_delta=np.array(rows["delta"]), sb_flag=np.array(rows["flag"]), sb_ewald=np.array(rows["ewald"]), sb_temp=np.array(rows["temp"]))


def reference_insertion_delta(box, ff, s, comp, rng):
    """One molecule of `comp` appended at a seeded place: Ewald_Total(after) - Ewald_Total(before) = Fourier delta minus the new
    molecule's self and intra-molecular exclusion (what GPU_EwaldDifference_General returns for an INSERTION,
    Ewald_Energy_Functions.h:547-576)."""
    if ff.no_charges:
        return {}
    ms = int(s.molsize[comp]); o = int(s.offsets[comp]); nm = int(s.natoms[comp]) // ms
    if (nm + 1) * ms > int(s.alloc[comp]):
        return {}
    L = np.array([box.cell[0], box.cell[4], box.cell[8]])
    tmpl = s.pos[o:o + ms]
    new = rng.random(3) * L + (tmpl - tmpl[0]) @ _random_rotation(rng).T
    E0, _, _ = orc.ref_ewald_total(box, s)
    E1, sa1, _ = orc.ref_ewald_total(box, _with_molecule(s, comp, nm, new))
    return dict(ins_pos=new, ins_ewald=np.array([E1[0] - E0[0], E1[2] - E0[2]]), ins_temp=sa1)


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
    # ---- moves with the reference's answers (pins the oracle's delta functions and, on the GPU, the move kernels)
    out.update(reference_move_deltas(box, ff, s, np.random.default_rng(seed + 5000)))
    out.update(reference_insertion_delta(box, ff, s, comp, np.random.default_rng(seed + 6000)))
    if deck.get("pockets"):
        pc, pr, inv = deck["pockets"][comp - s.nhost]
        out.update(pocket_centers=pc, pocket_radii=pr, pocket_invert=inv)
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
    # C: CO2 in NaX: 55 movable Na+ as framework component 1, cubic cell, shifted LJ with pair overrides, block pockets; 16 CO2
    save_config("C", add_adsorbates(load_deck_via_host(f"{EX}/CO2_NaX_Zeolite"), 2, 16, 80), 2, 1238)
    # D: Xe/Kr mixture: two monatomic adsorbates, no charges, O-O tail correction; 12 Kr + 12 Xe
    d = load_deck(f"{EX}/XeKr-Mixture")
    d = add_adsorbates(d, 1, 12, 78); d = add_adsorbates(d, 2, 12, 79)
    save_config("D", d, 2, 1237)
    np.savez_compressed(os.path.join(HERE, "rng_seed0.npz"), u=orc.ref_uniform_stream(0, 4096))


if __name__ == "__main__":
    main()
