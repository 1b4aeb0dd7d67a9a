"""TEST INFRASTRUCTURE: minimal reader of a gRASPA example deck (simulation.input,
force_field_mixing_rules.def, pseudo_atoms.def, <molecule>.def, <framework>.cif) into the
engine's host containers.  It follows the reference's parsers closely enough to reproduce
their numbers (read_data.cpp: ForceFieldParser :772-834, ForceField_Processing :1179-1247,
PseudoAtomParser :1320-1396, ReadFramework :1480-1760, MoleculeDefinitionParser :2044-2160,
read_Ewald_Parameters_from_input :609-702) and is used by tests/golden/make_golden.py only;
the GPU-side tests read the committed .npz fixtures instead.
"""
from __future__ import annotations

import os
import re

import numpy as np

from graspa_b200.types import Box, ForceField, System, beta_from_temperature
from oracle import oracle as orc


def _terms(line):
    return line.replace(",", " ").split()


def read_simulation_input(path):
    kv = {}
    comps = []
    with open(path) as f:
        for line in f:
            t = _terms(line)
            if not t or t[0].startswith("#"):
                continue
            if t[0] == "Component":
                comps.append({"name": t[3]})
            elif comps and t[0] in ("IdealGasRosenbluthWeight", "FugacityCoefficient", "WidomProbability", "TranslationProbability",
                                    "RotationProbability", "ReinsertionProbability", "SwapProbability", "IdentityChangeProbability",
                                    "CreateNumberOfMolecules", "MolFraction", "BlockPocketsFileName", "BlockPockets"):
                comps[-1][t[0]] = t[1]
            else:
                kv[t[0]] = t[1:]
    kv["components"] = comps
    return kv


def read_ff_mixing(path):
    names, eps, sig = [], [], []
    with open(path) as f:
        lines = f.read().splitlines()
    shifted = _terms(lines[1])[0] == "shifted"
    tail = _terms(lines[3])[0] == "yes"
    n = int(_terms(lines[5])[0])
    for i in range(7, 7 + n):
        t = _terms(lines[i])
        names.append(t[0]); eps.append(float(t[2])); sig.append(float(t[3]))
    return names, np.array(eps), np.array(sig), shifted, tail


def read_pseudo_atoms(path):
    with open(path) as f:
        lines = f.read().splitlines()
    n = int(_terms(lines[1])[0])
    names, mass, charge = [], [], []
    for i in range(3, 3 + n):
        t = _terms(lines[i])
        names.append(t[0]); mass.append(float(t[5])); charge.append(float(t[6]))
    return names, np.array(mass), np.array(charge)


def read_molecule(path, names, charges):
    with open(path) as f:
        lines = f.read().splitlines()
    ms = int(_terms(lines[5])[0])
    pos, typ = [], []
    for i in range(13, 13 + ms):
        t = _terms(lines[i])
        typ.append(names.index(t[1]))
        pos.append([float(x) for x in t[2:5]] if len(t) == 5 else [0.0, 0.0, 0.0])
    typ = np.array(typ, dtype=np.int64)
    return np.array(pos), typ, charges[typ]


def read_cif(path, names, pseudo_charge, unitcells, use_cif_charge):
    with open(path) as f:
        lines = f.read().splitlines()
    vals = {}
    for ln in lines:
        for key in ("_cell_length_a", "_cell_length_b", "_cell_length_c", "_cell_angle_alpha", "_cell_angle_beta", "_cell_angle_gamma"):
            if key in ln:
                vals[key] = float(_terms(ln)[1])
    cell = orc.cell_from_cif(vals["_cell_length_a"], vals["_cell_length_b"], vals["_cell_length_c"],
                             vals["_cell_angle_alpha"], vals["_cell_angle_beta"], vals["_cell_angle_gamma"], unitcells)
    cols = {}; count = 0; last = None
    for idx, ln in enumerate(lines):
        if "_atom_site" in ln:
            for k, key in enumerate(("_atom_site_label", "_atom_site_fract_x", "_atom_site_fract_y", "_atom_site_fract_z", "_atom_site_charge")):
                if key in ln:
                    cols[k] = count
            count += 1; last = idx
        elif last is not None:
            break
    fpos, typ, chg = [], [], []
    for ln in lines[last + 1:]:
        t = _terms(ln)
        if len(t) < 4:
            break
        label = re.sub(r"\d+$", "", t[cols[0]])
        j = names.index(label)
        fpos.append([float(t[cols[1]]), float(t[cols[2]]), float(t[cols[3]])])
        typ.append(j)
        chg.append(float(t[cols[4]]) if (use_cif_charge and 4 in cols) else pseudo_charge[j])
    fpos = np.array(fpos, dtype=np.float64).reshape(-1, 3); typ = np.array(typ, dtype=np.int64); chg = np.array(chg, dtype=np.float64)
    nx, ny, nz = unitcells
    shift = np.array([1.0 / nx, 1.0 / ny, 1.0 / nz])
    P, T, Q = [], [], []
    for ix in range(nx):
        for jy in range(ny):
            for kz in range(nz):
                sf = (fpos + np.array([ix, jy, kz], dtype=float)) * shift
                x = sf[:, 0] * cell[0] + sf[:, 1] * cell[3] + sf[:, 2] * cell[6]
                y = sf[:, 0] * cell[1] + sf[:, 1] * cell[4] + sf[:, 2] * cell[7]
                z = sf[:, 0] * cell[2] + sf[:, 1] * cell[5] + sf[:, 2] * cell[8]
                P.append(np.stack([x, y, z], axis=1)); T.append(typ); Q.append(chg)
    return cell, np.concatenate(P), np.concatenate(T), np.concatenate(Q)


def apply_tail_overrides(ffdef, names, eps, sig, shifted, cut_vdw, ut, te):
    """OverWriteTailCorrection, read_data.cpp:1132-1176 ("I J truncated yes" under "# rules to overwrite"); edits ut/te in place"""
    if not os.path.exists(ffdef):
        return
    with open(ffdef) as f:
        lines = f.read().splitlines()
    nover = int(_terms(lines[1])[0])
    n = len(names)
    for ln in lines[3:3 + nover]:
        t = _terms(ln)
        if len(t) == 4 and t[3] == "yes":
            i, j = names.index(t[0]), names.index(t[1])
            _, _, _, _, te1 = orc.ff_mix(eps, sig, [shifted] * n, [True] * n, cut_vdw)
            ut[i * n + j] = 1; ut[j * n + i] = 1; te[i * n + j] = te1[i * n + j]; te[j * n + i] = te1[i * n + j]


def lammps_recip_cutoff(cell, kmax):
    """ReciprocalCutOff of the LAMMPS-style Ewald set-up, read_data.cpp:669-685 (the tilt factors are zeroed there)"""
    lx, ly, lz = cell[0], cell[4], cell[8]
    ux, vy, wz = 2 * np.pi / lx, 2 * np.pi / ly, 2 * np.pi / lz
    kx = kmax[0] * ux; ky = kmax[0] * 0.0 + kmax[1] * vy; kz = kmax[0] * 0.0 + kmax[1] * 0.0 + kmax[2] * wz
    return max(kx * kx, ky * ky, kz * kz) * 1.00001


def read_restart_positions(path, adsorbate_index, molsize, box=None, with_charge=False):
    """RASPA-2 restart file -> positions (n, 3) [and charges] of one adsorbate component, RestartFileParser
    read_data.cpp:3000-3221: the block starts two lines after "Component: <i>"; `interval` position lines, then velocity,
    force, charge, scaling blocks of the same length.  Atoms other than the first of a molecule are re-wrapped to the
    nearest image of the first one (:3147-3160) when `box` is given."""
    with open(path) as f:
        lines = f.read().splitlines()
    start = None; nmol = 0
    for k, ln in enumerate(lines):
        if ln.find(f"Component: {adsorbate_index}") == 0:
            nmol = int(_terms(ln)[3]); start = k + 2
            break
    if start is None or nmol == 0:
        return (np.zeros((0, 3)), np.zeros(0)) if with_charge else np.zeros((0, 3))
    interval = nmol * molsize
    pos = np.zeros((interval, 3)); chg = np.zeros(interval)
    first = None
    for a in range(interval):
        t = _terms(lines[start + a])
        assert t[0].startswith("Adsorbate-atom-position"), "Cannot find matching strings in the range for reading positions!"
        p = np.array([float(t[3]), float(t[4]), float(t[5])])
        if int(t[2]) == 0:
            first = p
        elif box is not None:
            v = np.ascontiguousarray(p - first)
            orc.pbc(v, box)
            p = first + v
        pos[a] = p
        chg[a] = float(_terms(lines[start + 3 * interval + a])[3])
    return (pos, chg) if with_charge else pos


def load_deck(folder, unitcells=None, extra_alloc=0):
    """-> dict(box, ff, system, beta, ntrials, norient, names, sim).  Rigid single-component framework decks
    (configs A, B, D, E); separated framework components (config C) are handled by the host library, not here."""
    sim = read_simulation_input(os.path.join(folder, "simulation.input"))
    names, eps, sig, shifted, tail = read_ff_mixing(os.path.join(folder, "force_field_mixing_rules.def"))
    pnames, mass, pcharge = read_pseudo_atoms(os.path.join(folder, "pseudo_atoms.def"))
    assert pnames == names, "pseudo_atoms.def must list the force-field names in order (read_data.cpp:1344)"
    cut_vdw = float(sim.get("CutOffVDW", [12.0])[0]); cut_coul = float(sim.get("CutOffCoulomb", [12.0])[0])
    e, s, sh, ut, te = orc.ff_mix(eps, sig, [shifted] * len(eps), [tail] * len(eps), cut_vdw)
    apply_tail_overrides(os.path.join(folder, "force_field.def"), names, eps, sig, shifted, cut_vdw, ut, te)
    charge_method = sim.get("ChargeMethod", ["None"])[0].lower()
    no_charges = charge_method != "ewald"
    ff = ForceField(e, s, sh, cut_vdw, cut_coul, overlap=float(sim.get("OverlapCriteria", [1e5])[0]), no_charges=no_charges,
                    vdw_real_bias=True, use_tail=ut, tail_energy=te)
    uc = unitcells if unitcells is not None else tuple(int(x) for x in sim["UnitCells"][1:4])
    use_cif_charge = sim.get("UseChargesFromCIFFile", ["no"])[0].lower() == "yes"
    cell, fpos, ftyp, fchg = read_cif(os.path.join(folder, sim["FrameworkName"][0] + ".cif"), names, pcharge, uc, use_cif_charge)
    box = Box(cell)
    if not no_charges:
        box = orc.ewald_setup(box, cut_coul, float(sim.get("EwaldPrecision", [1e-6])[0]))
    natoms = [len(fpos)]; molsize = [len(fpos)]; alloc = [len(fpos)]
    P = [fpos]; Q = [fchg]; T = [ftyp]; M = [np.zeros(len(fpos), dtype=np.int64)]
    for comp in sim["components"]:
        mp, mt, mq = read_molecule(os.path.join(folder, comp["name"] + ".def"), names, pcharge)
        ms = len(mp); a = ms + extra_alloc
        pp = np.zeros((a, 3)); pp[:ms] = mp
        tt = np.zeros(a, dtype=np.int64); tt[:ms] = mt
        qq = np.zeros(a); qq[:ms] = mq
        natoms.append(0); molsize.append(ms); alloc.append(a)
        P.append(pp); Q.append(qq); T.append(tt); M.append(np.zeros(a, dtype=np.int64))
    system = System(1, np.array(natoms), np.array(molsize), np.concatenate(P), np.concatenate(Q), np.concatenate(T),
                    np.concatenate(M), alloc=np.array(alloc))
    T_K = float(sim.get("Temperature", [300.0])[0])
    beta = beta_from_temperature(T_K)
    return dict(box=box, ff=ff, system=system, beta=beta, temperature=T_K, names=names, sim=sim, mass=mass,
                ntrials=int(sim.get("NumberOfTrialPositions", [10])[0]), norient=int(sim.get("NumberOfTrialOrientations", [10])[0]))
