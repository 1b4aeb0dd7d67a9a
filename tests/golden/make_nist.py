"""Generates tests/golden/nist_spce.npz from the reference's NIST SPC/E known-answer example (build container only).

    python -m tests.golden.make_nist

Source: /root/reference/Examples/Reference_NIST_SPCE/Box-{1..4} -- 400/300/200/100 SPC/E waters in four triclinic
boxes, 10 A cutoffs, LAMMPS-style Ewald set-up (alpha 0.285, kmax 7 7 7), O-O tail correction switched on by
force_field.def.  Per box the fixture holds the inputs (cell, positions from RestartInitial/System_0/restartfile, force
field) and the reference's own printed energies (Box-N/output.txt, internal units, 5 decimals); the NIST table values of
readme.md:7-17 (Kelvin) are written into tests/test_oracle_golden.py.  The GPU-side tests read only the .npz.
"""
from __future__ import annotations

import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from graspa_b200.types import Box                      # noqa: E402
from oracle import oracle as orc                       # noqa: E402
from tests.support.raspa_inputs import (read_simulation_input, read_ff_mixing, read_pseudo_atoms, read_molecule,  # noqa: E402
                                        read_cif, read_restart_positions, lammps_recip_cutoff, apply_tail_overrides)

EX = "/root/reference/Examples/Reference_NIST_SPCE"


def printed(path):
    """the INITIAL STAGE block of the energy summary + the CPU Fourier line (fxn_main.h / ewald_preparation.h prints)"""
    txt = open(path).read()
    blk = txt[txt.index("*** INITIAL STAGE ***"):]
    out = {}
    for key, label in (("vdw_gg", r"VDW \[Guest-Guest\]:"), ("real_gg", r"Real Coulomb \[Guest-Guest\]:"),
                       ("ewald_gg", r"Ewald \[Guest-Guest\]:"), ("tail", r"Tail Correction Energy:"), ("total", r"Total Energy:")):
        out[key] = float(re.search(label + r"\s+(-?[\d.]+)", blk).group(1))
    out["fourier_gg"] = float(re.search(r"Guest-Guest Fourier: (-?[\d.]+)", txt).group(1))
    return out


def main():
    out = {}
    for b in (1, 2, 3, 4):
        d = f"{EX}/Box-{b}"
        sim = read_simulation_input(f"{d}/simulation.input")
        names, eps, sig, shifted, tail = read_ff_mixing(f"{d}/force_field_mixing_rules.def")
        pn, mass, pq = read_pseudo_atoms(f"{d}/pseudo_atoms.def")
        assert pn == names
        cv = float(sim["CutOffVDW"][0]); cc = float(sim["CutOffCoulomb"][0])
        e, s, sh, ut, te = orc.ff_mix(eps, sig, [shifted] * len(eps), [tail] * len(eps), cv)
        apply_tail_overrides(f"{d}/force_field.def", names, eps, sig, shifted, cv, ut, te)
        cell, fpos, _, _ = read_cif(f"{d}/Box-{b}.cif", names, pq, (1, 1, 1), False)
        assert len(fpos) == 0                                   # an empty "framework": the box only
        assert sim["Ewald_UseLAMMPS_Setup"][0].lower() == "yes"
        alpha = float(sim["Ewald_Alpha"][0]); kmax = [int(x) for x in sim["Ewald_kvectors"][:3]]
        box = Box(cell, alpha=alpha, kmax=tuple(kmax), use_lammps_ewald=True)
        rc = lammps_recip_cutoff(box.cell, kmax)
        mp, mt, mq = read_molecule(f"{d}/SPCE.def", names, pq)
        pos, chg = read_restart_positions(f"{d}/RestartInitial/System_0/restartfile", 0, len(mp), box, with_charge=True)
        assert np.allclose(chg, np.tile(mq, pos.shape[0] // len(mp)))
        nmol = pos.shape[0] // len(mp)
        ref = printed(f"{d}/output.txt")
        out.update({f"b{b}_cell": box.cell, f"b{b}_alpha": alpha, f"b{b}_kmax": np.array(kmax), f"b{b}_recip_cutoff": rc,
                    f"b{b}_pos": pos, f"b{b}_nmol": nmol,
                    f"b{b}_printed": np.array([ref[k] for k in ("vdw_gg", "real_gg", "ewald_gg", "tail", "total", "fourier_gg")])})
        print(f"Box-{b}: {nmol} molecules, recip cutoff {rc:.5f}, printed {ref}")
    out.update(eps=e, sigma=s, shift=sh, use_tail=ut, tail_energy=te, cutoff_vdw=cv, cutoff_coul=cc,
               overlap=float(sim["OverlapCriteria"][0]), mol_type=mt, mol_charge=mq, mol_pos=mp)
    np.savez_compressed(os.path.join(HERE, "nist_spce.npz"), **out)
    print(f"nist_spce.npz {os.path.getsize(os.path.join(HERE, 'nist_spce.npz')) / 1024:.0f} KB")


if __name__ == "__main__":
    main()
