"""TEST INFRASTRUCTURE ONLY: ctypes binding of the plain-C oracle (oracle/liboracle.so) and,
when present, of the reference-backed harness (oracle/_ref/libgraspa_ref_host.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from graspa_b200.types import Box, ForceField, System, TrialAtoms, species_counts, pseudo_atom_counts

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

f64p = C.POINTER(C.c_double)
i64p = C.POINTER(C.c_int64)
i32p = C.POINTER(C.c_int32)


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


class OrcBox(C.Structure):
    _fields_ = [("cell", C.c_double * 9), ("inv", C.c_double * 9), ("volume", C.c_double), ("alpha", C.c_double),
                ("prefactor", C.c_double), ("recip_cutoff", C.c_double), ("kmax", C.c_int32 * 3),
                ("cubic", C.c_int32), ("use_lammps_ewald", C.c_int32), ("pad_", C.c_int32)]


class OrcFF(C.Structure):
    _fields_ = [("epsilon", f64p), ("sigma", f64p), ("z", f64p), ("shift", f64p), ("c10", f64p),
                ("cutoff_vdw_sq", C.c_double), ("cutoff_coul_sq", C.c_double), ("overlap", C.c_double),
                ("ntypes", C.c_int32), ("no_charges", C.c_int32), ("vdw_real_bias", C.c_int32), ("use1264", C.c_int32)]


class OrcSystem(C.Structure):
    _fields_ = [("natoms", i64p), ("molsize", i64p), ("alloc", i64p), ("pos", f64p), ("scale", f64p), ("charge", f64p),
                ("scale_coul", f64p), ("type", i64p), ("molid", i64p), ("ncomp", C.c_int32), ("nhost", C.c_int32)]


class OrcAtoms(C.Structure):
    _fields_ = [("pos", f64p), ("scale", f64p), ("charge", f64p), ("scale_coul", f64p), ("type", i64p), ("n", C.c_int64)]


class OrcWidomCfg(C.Structure):
    _fields_ = [("sf_ads", f64p), ("sf_fw", f64p), ("excl_intra", C.c_double), ("excl_self", C.c_double),
                ("beta", C.c_double), ("ntrials", C.c_int32), ("norient", C.c_int32), ("comp", C.c_int32),
                ("has_charge", C.c_int32), ("ntypes", C.c_int32), ("has_tail", C.c_int32),
                ("npseudo", i64p), ("use_tail", i32p), ("tail_e", f64p), ("species_counts", i32p)]


def build(force=False):
    so = os.path.join(HERE, "liboracle.so")
    src = os.path.join(HERE, "graspa_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_coulomb_real.restype = C.c_double
        _LIB.orc_coulomb_real.argtypes = [C.c_double] * 6
        _LIB.orc_tail_total.restype = C.c_double
        _LIB.orc_tail_difference.restype = C.c_double
        _LIB.orc_tail_identity_swap.restype = C.c_double
        _LIB.orc_nvec.restype = C.c_int64
    return _LIB


def ref_available():
    return os.path.exists(os.path.join(HERE, "_ref", "libgraspa_ref_host.so"))


def ref():
    global _REF
    if _REF is None:
        _REF = C.CDLL(os.path.join(HERE, "_ref", "libgraspa_ref_host.so"))
        _REF.ref_tail_total.restype = C.c_double
        _REF.ref_tail_difference.restype = C.c_double
        _REF.ref_tail_identity_swap.restype = C.c_double
    return _REF


# ----------------------------------------------------------------- struct marshalling
def c_box(b: Box) -> OrcBox:
    o = OrcBox()
    o.cell[:] = list(b.cell); o.inv[:] = list(b.inv)
    o.volume = b.volume; o.alpha = b.alpha; o.prefactor = b.prefactor; o.recip_cutoff = b.recip_cutoff
    o.kmax[:] = list(b.kmax); o.cubic = int(b.cubic); o.use_lammps_ewald = int(b.use_lammps_ewald)
    return o


def c_ff(f: ForceField) -> OrcFF:
    o = OrcFF(_p(f.epsilon, f64p), _p(f.sigma, f64p), _p(f.z, f64p), _p(f.shift, f64p), _p(f.c10, f64p),
              f.cutoff_vdw_sq, f.cutoff_coul_sq, f.overlap, f.ntypes, int(f.no_charges), int(f.vdw_real_bias), int(f.use1264))
    o._keep = f
    return o


def c_sys(s: System) -> OrcSystem:
    o = OrcSystem(_p(s.natoms, i64p), _p(s.molsize, i64p), _p(s.alloc, i64p), _p(s.pos, f64p), _p(s.scale, f64p),
                  _p(s.charge, f64p), _p(s.scale_coul, f64p), _p(s.type, i64p), _p(s.molid, i64p), s.ncomp, s.nhost)
    o._keep = s
    return o


def c_atoms(t: TrialAtoms) -> OrcAtoms:
    o = OrcAtoms(_p(t.pos, f64p), _p(t.scale, f64p), _p(t.charge, f64p), _p(t.scale_coul, f64p), _p(t.type, i64p), t.n)
    o._keep = t
    return o


# ----------------------------------------------------------------- oracle calls
def ewald_setup(box: Box, cutoff_coul: float, precision: float) -> Box:
    ob = c_box(box)
    lib().orc_ewald_setup(C.c_double(cutoff_coul), C.c_double(precision), C.byref(ob))
    return Box(box.cell, alpha=ob.alpha, kmax=tuple(ob.kmax), recip_cutoff=ob.recip_cutoff, prefactor=ob.prefactor)


def cell_from_cif(a, b, c, al, be, ga, n):
    cell = np.zeros(9)
    lib().orc_cell_from_cif(*[C.c_double(x) for x in (a, b, c, al, be, ga)], C.c_int(n[0]), C.c_int(n[1]), C.c_int(n[2]), _p(cell, f64p))
    return cell


def pbc(v, box: Box):
    """minimum image of one difference vector, in place (maths.cuh:427-450)"""
    ob = c_box(box)
    lib().orc_pbc(_p(v, f64p), C.byref(ob))
    return v


def blocked_pocket(box: Box, centers, radii, pos, invert=False):
    """BlockedPocket (read_data.cpp:3466-3640) for one Cartesian position"""
    pk = np.ascontiguousarray(np.concatenate([np.asarray(centers, dtype=np.float64).reshape(-1, 3), np.asarray(radii, dtype=np.float64).reshape(-1, 1)], axis=1))
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    ob = c_box(box)
    return bool(lib().orc_blocked_pocket(C.byref(ob), _p(pk, f64p), C.c_int(pk.shape[0]), C.c_int(int(invert)), _p(pos, f64p)))


def ff_mix(eps, sig, shifted, tail, cutoff_vdw):
    n = len(eps)
    eps = np.ascontiguousarray(eps, dtype=np.float64); sig = np.ascontiguousarray(sig, dtype=np.float64)
    sh = np.ascontiguousarray(shifted, dtype=np.int32); tl = np.ascontiguousarray(tail, dtype=np.int32)
    e = np.zeros(n * n); s = np.zeros(n * n); shf = np.zeros(n * n); ut = np.zeros(n * n, dtype=np.int32); te = np.zeros(n * n)
    lib().orc_ff_mix(C.c_int(n), _p(eps, f64p), _p(sig, f64p), _p(sh, i32p), _p(tl, i32p), C.c_double(cutoff_vdw * cutoff_vdw),
                     _p(e, f64p), _p(s, f64p), _p(shf, f64p), _p(ut, i32p), _p(te, f64p))
    return e, s, shf, ut, te


def trial_positions(box, sys_, movetype, comp, start, ntrials, rnd, scale=1.0, scale_coul=1.0) -> TrialAtoms:
    rnd = np.ascontiguousarray(rnd, dtype=np.float64)
    pos = np.zeros((ntrials, 3)); sc = np.zeros(ntrials); q = np.zeros(ntrials); scc = np.zeros(ntrials); ty = np.zeros(ntrials, dtype=np.int64)
    ob, os_ = c_box(box), c_sys(sys_)
    lib().orc_trial_positions(C.byref(ob), C.byref(os_), C.c_int(movetype), C.c_int(comp), C.c_int64(start), C.c_int(ntrials),
                              _p(rnd, f64p), C.c_double(scale), C.c_double(scale_coul), _p(pos, f64p), _p(sc, f64p), _p(q, f64p), _p(scc, f64p), _p(ty, i64p))
    return TrialAtoms(pos, q, ty, sc, scc)


def trial_orientations(sys_, movetype, comp, start, chainsize, norient, rnd, fb_pos, fb_scale=1.0, fb_scale_coul=1.0) -> TrialAtoms:
    rnd = np.ascontiguousarray(rnd, dtype=np.float64); fb = np.ascontiguousarray(fb_pos, dtype=np.float64)
    n = norient * chainsize
    pos = np.zeros((n, 3)); sc = np.zeros(n); q = np.zeros(n); scc = np.zeros(n); ty = np.zeros(n, dtype=np.int64)
    os_ = c_sys(sys_)
    lib().orc_trial_orientations(C.byref(os_), C.c_int(movetype), C.c_int(comp), C.c_int64(start), C.c_int(chainsize), C.c_int(norient),
                                 _p(rnd, f64p), _p(fb, f64p), C.c_double(fb_scale), C.c_double(fb_scale_coul),
                                 _p(pos, f64p), _p(sc, f64p), _p(q, f64p), _p(scc, f64p), _p(ty, i64p))
    return TrialAtoms(pos, q, ty, sc, scc)


def trial_energies(box, ff, sys_, ntrials, chainsize, trial: TrialAtoms, new_comp, new_molid, excl_comp=-1, excl_mol=-1):
    out = np.zeros((ntrials, 4)); flag = np.zeros(ntrials, dtype=np.int32); counts = np.zeros(4, dtype=np.int64)
    ob, of, os_, ot = c_box(box), c_ff(ff), c_sys(sys_), c_atoms(trial)
    lib().orc_trial_energies(C.byref(ob), C.byref(of), C.byref(os_), C.c_int(ntrials), C.c_int(chainsize), C.byref(ot),
                             C.c_int(new_comp), C.c_int64(new_molid), C.c_int(excl_comp), C.c_int64(excl_mol),
                             _p(out, f64p), _p(flag, i32p), _p(counts, i64p))
    return out, flag, counts


def cbmc_finish(movetype, is_chain, energies, flags, beta, norm, uniform, vdw_real_bias=True, stored_in=0.0):
    """Host_sum_Widom_HGGG_SEPARATE + CBMC_FirstBead_Finish / chain tail on per-trial energies (n,4) and flags (n,).
    -> dict(success, selected (index among all trials), rosenbluth, stored_r, nsurv)"""
    energies = np.asarray(energies, dtype=np.float64); flags = np.asarray(flags)
    idx = [t for t in range(len(flags)) if not flags[t]]
    tot = energies[idx, 0] + energies[idx, 2]
    if vdw_real_bias:
        tot = tot + energies[idx, 1] + energies[idx, 3]
    rosen = np.ascontiguousarray(-beta * tot, dtype=np.float64)
    if rosen.size == 0:
        rosen = np.zeros(1)
    sel = C.c_int(0); R = C.c_double(0.0); stored = C.c_double(0.0)
    ok = lib().orc_cbmc_finish(C.c_int(movetype), C.c_int(int(is_chain)), _p(rosen, f64p), C.c_int(len(idx)), C.c_int(norm), C.c_double(uniform),
                               C.c_double(stored_in), C.byref(stored), C.byref(sel), C.byref(R))
    W = R.value
    real_sel = idx[sel.value] if (ok and idx) else 0
    if ok and idx and not vdw_real_bias:
        W *= np.exp(-beta * (energies[real_sel, 1] + energies[real_sel, 3]))
    return dict(success=bool(ok), selected=real_sel, rosenbluth=W, stored_r=stored.value, nsurv=len(idx))


def ewald_delta(box, pos, charge, scale_coul, nold, nnew, same_sf, cross_sf, want_temp=True):
    pos = np.ascontiguousarray(pos, dtype=np.float64); charge = np.ascontiguousarray(charge, dtype=np.float64)
    scale_coul = np.ascontiguousarray(scale_coul, dtype=np.float64)
    temp = np.zeros(2 * box.nvec) if want_temp else None
    out = np.zeros(2); act = C.c_int64(0)
    ob = c_box(box)
    lib().orc_ewald_delta(C.byref(ob), _p(pos, f64p), _p(charge, f64p), _p(scale_coul, f64p), C.c_int(nold), C.c_int(nnew),
                          _p(same_sf, f64p), _p(cross_sf, f64p), _p(temp, f64p), _p(out, f64p), C.byref(act))
    return out, temp, act.value


def ewald_total(box, sys_, no_charges=False):
    E = np.zeros(3); sa = np.zeros(2 * box.nvec); sf = np.zeros(2 * box.nvec)
    ob, os_ = c_box(box), c_sys(sys_)
    lib().orc_ewald_total(C.byref(ob), C.byref(os_), C.c_int(int(no_charges)), _p(E, f64p), _p(sa, f64p), _p(sf, f64p))
    return E, sa, sf


def exclusion_rigid(box, pos, charge, scale_coul):
    pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3); n = pos.shape[0]
    charge = np.ascontiguousarray(charge, dtype=np.float64); scale_coul = np.ascontiguousarray(scale_coul, dtype=np.float64)
    a = C.c_double(0); b = C.c_double(0); ob = c_box(box)
    lib().orc_exclusion_rigid(C.byref(ob), C.c_int(n), _p(pos, f64p), _p(charge, f64p), _p(scale_coul, f64p), C.byref(a), C.byref(b))
    return a.value, b.value


def tail_total(ff: ForceField, npseudo, volume):
    npseudo = np.ascontiguousarray(npseudo, dtype=np.int64)
    return lib().orc_tail_total(C.c_int(ff.ntypes), _p(npseudo, i64p), _p(ff.use_tail, i32p), _p(ff.tail_energy, f64p), C.c_double(volume))


def tail_difference(ff: ForceField, npseudo, volume, counts, sign):
    npseudo = np.ascontiguousarray(npseudo, dtype=np.int64); counts = np.ascontiguousarray(counts, dtype=np.int32)
    return lib().orc_tail_difference(C.c_int(ff.ntypes), _p(npseudo, i64p), _p(ff.use_tail, i32p), _p(ff.tail_energy, f64p),
                                     C.c_double(volume), _p(counts, i32p), C.c_int(sign))


def tail_identity_swap(ff: ForceField, npseudo, volume, new_counts, old_counts):
    npseudo = np.ascontiguousarray(npseudo, dtype=np.int64)
    nc = np.ascontiguousarray(new_counts, dtype=np.int32); oc = np.ascontiguousarray(old_counts, dtype=np.int32)
    return lib().orc_tail_identity_swap(C.c_int(ff.ntypes), _p(npseudo, i64p), _p(ff.use_tail, i32p), _p(ff.tail_energy, f64p),
                                        C.c_double(volume), _p(nc, i32p), _p(oc, i32p))


def single_body_delta(box, ff, sys_, comp, molid, old: TrialAtoms, new: TrialAtoms, do_new=True, do_old=True):
    out = np.zeros(6); flag = C.c_int32(0)
    ob, of, os_, oo, on = c_box(box), c_ff(ff), c_sys(sys_), c_atoms(old), c_atoms(new)
    lib().orc_single_body_delta(C.byref(ob), C.byref(of), C.byref(os_), C.c_int(comp), C.c_int64(molid), C.byref(oo), C.byref(on),
                                C.c_int(int(do_new)), C.c_int(int(do_old)), _p(out, f64p), C.byref(flag))
    return out, flag.value


def total_vdw_real(box, ff, sys_):
    out = np.zeros(6); ob, of, os_ = c_box(box), c_ff(ff), c_sys(sys_)
    lib().orc_total_vdw_real(C.byref(ob), C.byref(of), C.byref(os_), _p(out, f64p))
    return out


class WidomSetup:
    """Everything a Widom / CBMC insertion of component ``comp`` needs besides the randoms."""

    def __init__(self, box, ff, sys_, comp, beta, ntrials, norient, sf_ads=None, sf_fw=None):
        self.box, self.ff, self.sys, self.comp, self.beta = box, ff, sys_, comp, beta
        self.ntrials, self.norient = ntrials, norient
        o = int(sys_.offsets[comp]); ms = int(sys_.molsize[comp])
        self.molsize = ms
        self.has_charge = bool(np.any(np.abs(sys_.charge[o:o + ms]) > 1e-10)) and not ff.no_charges
        if sf_ads is None and not ff.no_charges:
            _, sf_ads, sf_fw = ewald_total(box, sys_)
        self.sf_ads = np.ascontiguousarray(sf_ads if sf_ads is not None else np.zeros(2 * max(box.nvec, 1)))
        self.sf_fw = np.ascontiguousarray(sf_fw if sf_fw is not None else np.zeros(2 * max(box.nvec, 1)))
        if ff.no_charges:
            self.excl = (0.0, 0.0)
        else:
            self.excl = exclusion_rigid(box, sys_.pos[o:o + ms], sys_.charge[o:o + ms], sys_.scale_coul[o:o + ms])
        self.npseudo = pseudo_atom_counts(sys_, ff.ntypes)
        self.counts = species_counts(sys_, comp, ff.ntypes)

    def cfg(self) -> OrcWidomCfg:
        ff = self.ff
        c = OrcWidomCfg(_p(self.sf_ads, f64p), _p(self.sf_fw, f64p), self.excl[0], self.excl[1], self.beta,
                        self.ntrials, self.norient, self.comp, int(self.has_charge), ff.ntypes, int(ff.has_tail),
                        _p(self.npseudo, i64p), _p(ff.use_tail, i32p), _p(ff.tail_energy, f64p), _p(self.counts, i32p))
        c._keep = self
        return c


def widom_insertion(ws: WidomSetup, rnd_fb, u_fb, rnd_or, u_or):
    rnd_fb = np.ascontiguousarray(rnd_fb, dtype=np.float64); rnd_or = np.ascontiguousarray(rnd_or, dtype=np.float64)
    out = np.zeros(8); stage = C.c_int32(0); sel = np.zeros(2, dtype=np.int32); pos = np.zeros((ws.molsize, 3))
    counts = np.zeros(5, dtype=np.int64)
    ob, of, os_, cfg = c_box(ws.box), c_ff(ws.ff), c_sys(ws.sys), ws.cfg()
    lib().orc_widom_insertion(C.byref(ob), C.byref(of), C.byref(os_), C.byref(cfg), _p(rnd_fb, f64p), C.c_double(u_fb),
                              _p(rnd_or, f64p), C.c_double(u_or), _p(out, f64p), C.byref(stage), _p(sel, i32p), _p(pos, f64p), _p(counts, i64p))
    return out, stage.value, sel, pos, counts


def widom_batch(ws: WidomSetup, rnd, uni, nthreads=0):
    """rnd: (n, ntrials+norient, 3); uni: (n, 2)"""
    rnd = np.ascontiguousarray(rnd, dtype=np.float64); uni = np.ascontiguousarray(uni, dtype=np.float64)
    n = uni.shape[0]
    out = np.zeros((n, 8)); stage = np.zeros(n, dtype=np.int32); counts = np.zeros(5, dtype=np.int64)
    ob, of, os_, cfg = c_box(ws.box), c_ff(ws.ff), c_sys(ws.sys), ws.cfg()
    lib().orc_widom_batch(C.byref(ob), C.byref(of), C.byref(os_), C.byref(cfg), C.c_int64(n), _p(rnd, f64p), _p(uni, f64p),
                          C.c_int(nthreads), _p(out, f64p), _p(stage, i32p), _p(counts, i64p))
    return out, stage, counts


def max_threads():
    return int(lib().orc_max_threads())


def uniform_stream(seed, n):
    out = np.zeros(n)
    lib().orc_uniform_stream(C.c_int(seed), C.c_int64(n), _p(out, f64p))
    return out


# ----------------------------------------------------------------- reference-backed harness calls
def ref_trial_energies(box, ff, sys_, ntrials, chainsize, trial, new_comp, new_molid, excl_comp=-1, excl_mol=-1):
    s = sys_.compact()
    out = np.zeros((ntrials, 4)); flag = np.zeros(ntrials, dtype=np.int32); counts = np.zeros(3, dtype=np.int64)
    ref().ref_trial_energies(C.c_int(s.ncomp), C.c_int(s.nhost), _p(s.natoms, i64p), _p(s.molsize, i64p), _p(s.pos, f64p), _p(s.scale, f64p),
                             _p(s.charge, f64p), _p(s.scale_coul, f64p), _p(s.type, i64p), _p(s.molid, i64p),
                             _p(box.cell, f64p), _p(box.inv, f64p), C.c_int(int(box.cubic)), C.c_double(box.prefactor), C.c_double(box.alpha),
                             C.c_int(ff.ntypes), _p(ff.epsilon, f64p), _p(ff.sigma, f64p), _p(ff.z, f64p), _p(ff.shift, f64p), _p(ff.c10, f64p),
                             C.c_double(ff.cutoff_vdw_sq), C.c_double(ff.cutoff_coul_sq), C.c_double(ff.overlap), C.c_int(int(ff.no_charges)), C.c_int(int(ff.use1264)),
                             C.c_int(ntrials), C.c_int(chainsize), _p(trial.pos, f64p), _p(trial.scale, f64p), _p(trial.charge, f64p),
                             _p(trial.scale_coul, f64p), _p(trial.type, i64p), C.c_longlong(new_molid), C.c_int(new_comp),
                             C.c_int(excl_comp), C.c_longlong(excl_mol), _p(out, f64p), _p(flag, i32p), _p(counts, i64p))
    return out, flag, counts


def ref_ewald_total(box, sys_, no_charges=False):
    s = sys_.compact()
    E = np.zeros(3); sa = np.zeros(2 * box.nvec); sf = np.zeros(2 * box.nvec)
    kmax = np.asarray(box.kmax, dtype=np.int32)
    ref().ref_ewald_total(C.c_int(s.ncomp), C.c_int(s.nhost), _p(s.natoms, i64p), _p(s.molsize, i64p), _p(s.pos, f64p), _p(s.scale, f64p),
                          _p(s.charge, f64p), _p(s.scale_coul, f64p), _p(s.type, i64p), _p(s.molid, i64p),
                          _p(box.cell, f64p), _p(box.inv, f64p), C.c_int(int(box.cubic)), C.c_double(box.volume), C.c_double(box.prefactor),
                          C.c_double(box.alpha), _p(kmax, i32p), C.c_double(box.recip_cutoff), C.c_int(int(box.use_lammps_ewald)),
                          C.c_int(int(no_charges)), _p(E, f64p), _p(sa, f64p), _p(sf, f64p))
    return E, sa, sf


def ref_exclusion_rigid(box, pos, charge, scale_coul):
    pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3); n = pos.shape[0]
    charge = np.ascontiguousarray(charge, dtype=np.float64); scale_coul = np.ascontiguousarray(scale_coul, dtype=np.float64)
    a = C.c_double(0); b = C.c_double(0)
    ref().ref_exclusion_rigid(C.c_int(n), _p(pos, f64p), _p(charge, f64p), _p(scale_coul, f64p), _p(box.cell, f64p), _p(box.inv, f64p),
                              C.c_int(int(box.cubic)), C.c_double(box.prefactor), C.c_double(box.alpha), C.byref(a), C.byref(b))
    return a.value, b.value


def ref_pair(box, ff, posA, posB, typeA, typeB, scaling, qA, qB, scaling_coul):
    row = typeA * ff.ntypes + typeB
    ffarg = np.array([ff.epsilon[row], ff.sigma[row], ff.z[row], ff.shift[row], ff.c10[row]])
    out = np.zeros(6)
    posA = np.ascontiguousarray(posA, dtype=np.float64); posB = np.ascontiguousarray(posB, dtype=np.float64)
    ref().ref_pair(_p(box.cell, f64p), _p(box.inv, f64p), C.c_int(int(box.cubic)), _p(posA, f64p), _p(posB, f64p), _p(ffarg, f64p),
                   C.c_double(scaling), C.c_int(int(ff.use1264)), C.c_double(qA), C.c_double(qB), C.c_double(scaling_coul),
                   C.c_double(box.prefactor), C.c_double(box.alpha), C.c_double(ff.cutoff_vdw_sq), C.c_double(ff.cutoff_coul_sq),
                   C.c_int(int(ff.no_charges)), _p(out, f64p))
    return out


def _species_lists(ncomp_counts):
    nent, types, cnts = [], [], []
    for counts in ncomp_counts:
        nz = [(t, int(c)) for t, c in enumerate(counts) if c > 0]
        nent.append(len(nz)); types += [t for t, _ in nz]; cnts += [c for _, c in nz]
    return (np.asarray(nent, dtype=np.int32), np.asarray(types + [0], dtype=np.int32), np.asarray(cnts + [0], dtype=np.int32))


def ref_tail_total(ff, npseudo, volume):
    npseudo = np.ascontiguousarray(npseudo, dtype=np.int64)
    return ref().ref_tail_total(C.c_int(ff.ntypes), _p(npseudo, i64p), _p(ff.use_tail, i32p), _p(ff.tail_energy, f64p), C.c_double(volume))


def ref_tail_difference(ff, npseudo, volume, counts_per_comp, comp, insertion=True):
    npseudo = np.ascontiguousarray(npseudo, dtype=np.int64)
    nent, ty, cn = _species_lists(counts_per_comp)
    mt = ref().ref_movetype_insertion() if insertion else ref().ref_movetype_deletion()
    return ref().ref_tail_difference(C.c_int(ff.ntypes), _p(npseudo, i64p), _p(ff.use_tail, i32p), _p(ff.tail_energy, f64p), C.c_double(volume),
                                     C.c_int(len(counts_per_comp)), _p(nent, i32p), _p(ty, i32p), _p(cn, i32p), C.c_int(comp), C.c_int(mt))


def ref_tail_identity_swap(ff, npseudo, volume, counts_per_comp, newcomp, oldcomp):
    npseudo = np.ascontiguousarray(npseudo, dtype=np.int64)
    nent, ty, cn = _species_lists(counts_per_comp)
    return ref().ref_tail_identity_swap(C.c_int(ff.ntypes), _p(npseudo, i64p), _p(ff.use_tail, i32p), _p(ff.tail_energy, f64p), C.c_double(volume),
                                        C.c_int(len(counts_per_comp)), _p(nent, i32p), _p(ty, i32p), _p(cn, i32p), C.c_int(newcomp), C.c_int(oldcomp))


def ref_uniform_stream(seed, n):
    out = np.zeros(n)
    ref().ref_uniform_stream(C.c_int(seed), C.c_longlong(n), _p(out, f64p))
    return out


def ref_inverse_matrix(cell):
    cell = np.ascontiguousarray(cell, dtype=np.float64); inv = np.zeros(9); det = C.c_double(0)
    ref().ref_inverse_matrix(_p(cell, f64p), _p(inv, f64p), C.byref(det))
    return inv, det.value


def scale_positions(box: Box, sys_: System, scale: float):
    """ScalePositions (mc_box.h:18-64) restated: every molecule of components >= 1 follows its first atom, which scales with the box;
    the other atoms keep their minimum-image offset (PBC, maths.cuh:427-450, old box) from it.  -> positions (n, 3) of the scaled state.
    The framework (component 0) is not scaled (ScaleFirstComponentFramework = false, mc_box.h:204)."""
    inv = box.inv.reshape(3, 3); cell = box.cell.reshape(3, 3)
    pos = sys_.pos.copy()
    for c in range(1, sys_.ncomp):
        o = int(sys_.offsets[c]); ms = int(sys_.molsize[c])
        for m in range(int(sys_.natoms[c]) // max(ms, 1)):
            first = sys_.pos[o + m * ms].copy()
            d = sys_.pos[o + m * ms:o + (m + 1) * ms] - first
            if box.cubic:
                L = np.array([cell[0, 0], cell[1, 1], cell[2, 2]]); iL = np.array([inv[0, 0], inv[1, 1], inv[2, 2]])
                d = d - np.trunc(d * iL + np.where(d >= 0.0, 0.5, -0.5)) * L
            else:
                f = d @ inv
                f = f - np.trunc(f + np.where(f >= 0.0, 0.5, -0.5))
                d = f @ cell
            pos[o + m * ms:o + (m + 1) * ms] = first * scale + d
    return pos
