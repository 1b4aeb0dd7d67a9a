"""ctypes binding of libgraspa_b200.so (include/graspa_b200.h).

This is plumbing for tests/ and bench.py: every call goes through the C ABI a host program
would bind, with host numpy buffers in and out.  There is no Python or CPU implementation of
any energy behind it -- if the CUDA library is missing or no GPU is present, it raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .types import Box, ForceField, System, TrialAtoms

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgraspa_b200.so")

f64p = C.POINTER(C.c_double)
i64p = C.POINTER(C.c_int64)
u64p = C.POINTER(C.c_uint64)
i32p = C.POINTER(C.c_int32)

# every symbol include/graspa_b200.h declares (tests check the library exports all of them)
DECLARED_SYMBOLS = [
    "gb_abi_version", "gb_last_error", "gb_engine_create", "gb_engine_destroy", "gb_device_info", "gb_synchronize", "gb_stream",
    "gb_upload_forcefield", "gb_upload_box", "gb_set_components", "gb_upload_atoms", "gb_download_atoms", "gb_snapshot_molecules",
    "gb_upload_structure_factors", "gb_download_structure_factors", "gb_set_exclusion_constants", "gb_upload_random_pool", "gb_set_block_pockets",
    "gb_set_cbmc", "gb_get_pseudo_atom_counts", "gb_cbmc_first_bead", "gb_cbmc_chain", "gb_cbmc_grown_positions", "gb_reinsertion_store", "gb_move_insertion", "gb_move_deletion", "gb_move_reinsertion", "gb_move_single_body", "gb_move_identity_swap",
    "gb_trial_energies", "gb_single_body_propose", "gb_single_body_delta", "gb_single_body_delta_explicit",
    "gb_ewald_delta", "gb_ewald_delta_identity_swap", "gb_ewald_delta_explicit", "gb_ewald_commit",
    "gb_lambda_change_delta", "gb_ewald_delta_lambda_change", "gb_accept_lambda_change", "gb_cbcf_set_scale", "gb_cbcf_deletion_stage",
    "gb_tail_total", "gb_tail_difference", "gb_tail_identity_swap",
    "gb_accept_translation", "gb_accept_insertion", "gb_accept_deletion", "gb_accept_reinsertion", "gb_accept_identity_swap", "gb_append_molecule",
    "gb_number_of_molecules", "gb_total_vdw_real", "gb_total_ewald", "gb_volume_move_trial", "gb_volume_move_finish",
    "gb_widom_batch", "gb_widom_first_bead_success",
    "gb_launch_count", "gb_timing_enable", "gb_timing_read", "gb_move_server", "gb_measure_fp64_peak",
]


class GbBox(C.Structure):
    _fields_ = [("cell", C.c_double * 9), ("inverse_cell", C.c_double * 9), ("volume", C.c_double), ("alpha", C.c_double),
                ("prefactor", C.c_double), ("reciprocal_cutoff", C.c_double), ("kmax", C.c_int32 * 3), ("cubic", C.c_int32),
                ("use_lammps_ewald", C.c_int32), ("reserved", C.c_int32)]


class GbForceField(C.Structure):
    _fields_ = [("epsilon", f64p), ("sigma", f64p), ("z", f64p), ("shift", f64p), ("c10", f64p),
                ("cutoff_vdw_sq", C.c_double), ("cutoff_coul_sq", C.c_double), ("overlap_criteria", C.c_double),
                ("size", C.c_int32), ("no_charges", C.c_int32), ("vdw_real_bias", C.c_int32), ("use1264", C.c_int32)]


class GbTailTable(C.Structure):
    _fields_ = [("use_tail", i32p), ("energy", f64p), ("size", C.c_int32), ("reserved", C.c_int32)]


class GbAtoms(C.Structure):
    _fields_ = [("pos", f64p), ("scale", f64p), ("charge", f64p), ("scale_coul", f64p), ("type", u64p), ("molid", u64p),
                ("n_upload", C.c_int64), ("n_live", C.c_int64), ("n_alloc", C.c_int64), ("molsize", C.c_int64)]


class GbMoveEnergy(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("storedHGVDW", "storedHGReal", "storedHGEwaldE", "HHVDW", "HGVDW", "GGVDW",
                                          "HHReal", "HGReal", "GGReal", "HHEwaldE", "HGEwaldE", "GGEwaldE", "TailE", "DNN_E")]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class GbCbmcResult(C.Structure):
    _fields_ = [("rosenbluth", C.c_double), ("stored_r", C.c_double), ("energy", C.c_double * 4), ("selected_pos", C.c_double * 3),
                ("success", C.c_int32), ("selected", C.c_int32), ("n_survivors", C.c_int32), ("reserved", C.c_int32)]


class GbMoveResult(C.Structure):
    _fields_ = [("first_bead", GbCbmcResult), ("chain", GbCbmcResult), ("old_first_bead", GbCbmcResult), ("old_chain", GbCbmcResult),
                ("ewald", C.c_double * 2), ("tail", C.c_double), ("delta", GbMoveEnergy), ("overlap", C.c_int32),
                ("uniforms_used", C.c_int32), ("pool_used", C.c_int32), ("success", C.c_int32)]

    def as_dict(self):
        d = {k: Engine._cbmc_dict(getattr(self, k), 0) for k in ("first_bead", "chain", "old_first_bead", "old_chain")}
        d.update(ewald=(self.ewald[0], self.ewald[1]), tail=self.tail, delta=self.delta.as_dict(), overlap=bool(self.overlap),
                 uniforms_used=self.uniforms_used, pool_used=self.pool_used, success=bool(self.success))
        return d


class GbWidomInputs(C.Structure):
    _fields_ = [("pool3", C.c_void_p), ("n_pool", C.c_int64), ("fb_index", C.c_void_p), ("or_index", C.c_void_p),
                ("uniforms", C.c_void_p), ("inputs_on_device", C.c_int32), ("n_blocks", C.c_int32),
                ("global_first", C.c_int64), ("global_n", C.c_int64), ("sums_device", C.c_void_p),
                ("resume_first_bead", C.c_int32), ("reserved", C.c_int32)]


class EngineError(RuntimeError):
    pass


_LIB = None


def build(force=False):
    """Compile the CUDA extension in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    src_dir = os.path.join(HERE, "csrc")
    srcs = [os.path.join(src_dir, f) for f in os.listdir(src_dir) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(HERE, "..", "include", "graspa_b200.h"))
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", src_dir], stdout=subprocess.DEVNULL)
    return LIB_PATH


def load_library(path=None):
    """dlopen the engine.  Fails loudly when the library has not been built: there is no fallback."""
    global _LIB
    if _LIB is None:
        p = path or os.environ.get("GRASPA_B200_LIB") or LIB_PATH      # the override serves kernel experiments (build_dbg/)
        if not os.path.exists(p):
            raise EngineError(f"{p} is missing: build it with graspa_b200.engine.build() / make -C graspa_b200/csrc; "
                              "graspa_b200 has no CPU or Python fallback")
        _LIB = C.CDLL(p)
        _LIB.gb_last_error.restype = C.c_char_p
        _LIB.gb_stream.restype = C.c_void_p
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


class Engine:
    """One engine per GPU.  Mirrors the C ABI one to one."""

    def __init__(self, device=-1):
        self.lib = load_library()
        self.h = C.c_void_p()
        self._chk(self.lib.gb_engine_create(C.byref(self.h), C.c_int(device)))
        self.system = None
        self._keep = []

    def _chk(self, rc):
        if rc != 0:
            raise EngineError(f"graspa_b200 error {rc}: {self.lib.gb_last_error().decode()}")

    def close(self):
        if self.h:
            self.lib.gb_engine_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------ set-up
    def device_info(self):
        sm, ma, mi, sh = C.c_int(), C.c_int(), C.c_int(), C.c_int64()
        self._chk(self.lib.gb_device_info(self.h, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(sh)))
        return dict(sm_count=sm.value, cc=(ma.value, mi.value), smem_optin=sh.value)

    def upload_forcefield(self, ff: ForceField):
        g = GbForceField(_p(ff.epsilon, f64p), _p(ff.sigma, f64p), _p(ff.z, f64p), _p(ff.shift, f64p), _p(ff.c10, f64p),
                         ff.cutoff_vdw_sq, ff.cutoff_coul_sq, ff.overlap, ff.ntypes, int(ff.no_charges), int(ff.vdw_real_bias), int(ff.use1264))
        t = GbTailTable(_p(ff.use_tail, i32p), _p(ff.tail_energy, f64p), ff.ntypes, 0)
        self._chk(self.lib.gb_upload_forcefield(self.h, C.byref(g), C.byref(t)))
        self.ff = ff

    @staticmethod
    def _gb_box(box: Box):
        g = GbBox()
        g.cell[:] = list(box.cell); g.inverse_cell[:] = list(box.inv)
        g.volume = box.volume; g.alpha = box.alpha; g.prefactor = box.prefactor; g.reciprocal_cutoff = box.recip_cutoff
        g.kmax[:] = list(box.kmax); g.cubic = int(box.cubic); g.use_lammps_ewald = int(box.use_lammps_ewald)
        return g

    def upload_box(self, box: Box):
        g = self._gb_box(box)
        self._chk(self.lib.gb_upload_box(self.h, C.byref(g)))
        self.box = box

    def volume_move_trial(self, new_box: Box, scale):
        """VolumeMove up to the acceptance test (mc_box.h:196-254) -> (energies of the scaled system, overlap flag)"""
        g = self._gb_box(new_box); m = GbMoveEnergy(); ov = C.c_int32(0)
        self._chk(self.lib.gb_volume_move_trial(self.h, C.byref(g), C.c_double(scale), C.byref(m), C.byref(ov)))
        self._pending_box = new_box
        return m.as_dict(), int(ov.value)

    def volume_move_finish(self, accept):
        self._chk(self.lib.gb_volume_move_finish(self.h, C.c_int32(int(bool(accept)))))
        if accept:
            self.box = self._pending_box

    def upload_system(self, s: System):
        self._chk(self.lib.gb_set_components(self.h, C.c_int32(s.ncomp), C.c_int32(s.nhost)))
        off = s.offsets
        for c in range(s.ncomp):
            sl = slice(int(off[c]), int(off[c + 1]))
            pos = np.ascontiguousarray(s.pos[sl]); sc = np.ascontiguousarray(s.scale[sl]); q = np.ascontiguousarray(s.charge[sl])
            scc = np.ascontiguousarray(s.scale_coul[sl]); ty = np.ascontiguousarray(s.type[sl].astype(np.uint64))
            mo = np.ascontiguousarray(s.molid[sl].astype(np.uint64))
            a = GbAtoms(_p(pos, f64p), _p(sc, f64p), _p(q, f64p), _p(scc, f64p), _p(ty, u64p), _p(mo, u64p),
                        int(s.alloc[c]), int(s.natoms[c]), int(s.alloc[c]), int(s.molsize[c]))
            self._chk(self.lib.gb_upload_atoms(self.h, C.c_int32(c), C.byref(a)))
        self.system = s

    def setup(self, box: Box, ff: ForceField, system: System, beta=None, ntrials=10, norient=10):
        self.upload_forcefield(ff)
        self.upload_box(box)
        self.upload_system(system)
        if beta is not None:
            self.set_cbmc(ntrials, norient, beta)
        return self

    def download_atoms(self, c):
        n = int(self.system.alloc[c])
        pos = np.zeros((n, 3)); sc = np.zeros(n); q = np.zeros(n); scc = np.zeros(n)
        ty = np.zeros(n, dtype=np.uint64); mo = np.zeros(n, dtype=np.uint64); nl = C.c_int64()
        self._chk(self.lib.gb_download_atoms(self.h, C.c_int32(c), _p(pos, f64p), _p(sc, f64p), _p(q, f64p), _p(scc, f64p),
                                             _p(ty, u64p), _p(mo, u64p), C.byref(nl)))
        return dict(pos=pos, scale=sc, charge=q, scale_coul=scc, type=ty.astype(np.int64), molid=mo.astype(np.int64), n_live=nl.value)

    def upload_structure_factors(self, ads, fw):
        ads = np.ascontiguousarray(ads, dtype=np.float64); fw = np.ascontiguousarray(fw, dtype=np.float64)
        self._chk(self.lib.gb_upload_structure_factors(self.h, _p(ads, f64p), _p(fw, f64p)))

    def download_structure_factors(self):
        n = 2 * self.box.nvec
        a = np.zeros(n); f = np.zeros(n); t = np.zeros(n)
        self._chk(self.lib.gb_download_structure_factors(self.h, _p(a, f64p), _p(f, f64p), _p(t, f64p)))
        return a, f, t

    def set_exclusion_constants(self, c, intra, atom, rigid=True, has_charge=True):
        self._chk(self.lib.gb_set_exclusion_constants(self.h, C.c_int32(c), C.c_double(intra), C.c_double(atom), C.c_int32(int(rigid)), C.c_int32(int(has_charge))))

    def set_cbmc(self, ntrials, norient, beta):
        self._chk(self.lib.gb_set_cbmc(self.h, C.c_int32(ntrials), C.c_int32(norient), C.c_double(beta)))
        self.ntrials, self.norient, self.beta = ntrials, norient, beta

    def pseudo_atom_counts(self):
        out = np.zeros(self.ff.ntypes, dtype=np.int64)
        self._chk(self.lib.gb_get_pseudo_atom_counts(self.h, _p(out, i64p)))
        return out

    # ------------------------------------------------------------ energies
    def trial_energies(self, ntrials, chainsize, trial: TrialAtoms, new_comp, new_molid, excl_comp=-1, excl_mol=-1):
        out = np.zeros((ntrials, 4)); flag = np.zeros(ntrials, dtype=np.int32)
        ty = np.ascontiguousarray(trial.type.astype(np.uint64))
        self._chk(self.lib.gb_trial_energies(self.h, C.c_int32(ntrials), C.c_int32(chainsize), _p(trial.pos, f64p), _p(trial.scale, f64p),
                                             _p(trial.charge, f64p), _p(trial.scale_coul, f64p), _p(ty, u64p), C.c_int32(new_comp),
                                             C.c_int64(new_molid), C.c_int32(excl_comp), C.c_int64(excl_mol), _p(out, f64p), _p(flag, i32p)))
        return out, flag

    def ewald_delta_explicit(self, framework_moved, nold, nnew, pos, charge, scale_coul):
        pos = np.ascontiguousarray(pos, dtype=np.float64); charge = np.ascontiguousarray(charge, dtype=np.float64)
        scale_coul = np.ascontiguousarray(scale_coul, dtype=np.float64)
        out = np.zeros(2)
        self._chk(self.lib.gb_ewald_delta_explicit(self.h, C.c_int32(int(framework_moved)), C.c_int32(nold), C.c_int32(nnew),
                                                   _p(pos, f64p), _p(charge, f64p), _p(scale_coul, f64p), _p(out, f64p)))
        return out

    def ewald_commit(self, c):
        self._chk(self.lib.gb_ewald_commit(self.h, C.c_int32(c)))

    def tail_total(self):
        o = C.c_double(); self._chk(self.lib.gb_tail_total(self.h, C.byref(o))); return o.value

    def tail_difference(self, c, move_type):
        o = C.c_double(); self._chk(self.lib.gb_tail_difference(self.h, C.c_int32(c), C.c_int32(move_type), C.byref(o))); return o.value

    def tail_identity_swap(self, newc, oldc):
        o = C.c_double(); self._chk(self.lib.gb_tail_identity_swap(self.h, C.c_int32(newc), C.c_int32(oldc), C.byref(o))); return o.value

    def total_vdw_real(self):
        m = GbMoveEnergy(); self._chk(self.lib.gb_total_vdw_real(self.h, C.byref(m))); return m.as_dict()

    def total_ewald(self, store=False):
        m = GbMoveEnergy(); self._chk(self.lib.gb_total_ewald(self.h, C.c_int32(int(store)), C.byref(m))); return m.as_dict()

    # ------------------------------------------------------------ single-move path
    def snapshot_molecules(self, comp, first, count):
        n = int(count) * int(self.system.molsize[comp])
        pos = np.zeros((n, 3)); q = np.zeros(n); sc = np.zeros(n); scc = np.zeros(n)
        self._chk(self.lib.gb_snapshot_molecules(self.h, C.c_int32(comp), C.c_int64(first), C.c_int64(count), _p(pos, f64p), _p(q, f64p), _p(sc, f64p), _p(scc, f64p)))
        return dict(pos=pos, charge=q, scale=sc, scale_coul=scc)

    def set_block_pockets(self, comp, centers, radii, invert=False):
        centers = np.ascontiguousarray(centers, dtype=np.float64).reshape(-1, 3); radii = np.ascontiguousarray(radii, dtype=np.float64)
        self._chk(self.lib.gb_set_block_pockets(self.h, C.c_int32(comp), C.c_int32(len(radii)), _p(centers, f64p), _p(radii, f64p), C.c_int32(int(invert))))

    def upload_random_pool(self, rnd3):
        rnd3 = np.ascontiguousarray(rnd3, dtype=np.float64).reshape(-1, 3)
        self._chk(self.lib.gb_upload_random_pool(self.h, _p(rnd3, f64p), C.c_int64(rnd3.shape[0])))

    @staticmethod
    def _cbmc_dict(r: GbCbmcResult, used):
        return dict(rosenbluth=r.rosenbluth, stored_r=r.stored_r, energy=np.array(list(r.energy)), selected_pos=np.array(list(r.selected_pos)),
                    success=bool(r.success), selected=int(r.selected), n_survivors=int(r.n_survivors), uniform_used=int(used))

    def cbmc_first_bead(self, cbmc_type, comp, molecule, pool_offset, uniform, scale=(1.0, 1.0), stored_r=0.0,
                        excl_comp=-1, excl_mol=-1, preset_pos=None):
        sc = (C.c_double * 2)(*scale); r = GbCbmcResult(); used = C.c_int32()
        pp = (C.c_double * 3)(*preset_pos) if preset_pos is not None else None
        self._chk(self.lib.gb_cbmc_first_bead(self.h, C.c_int32(cbmc_type), C.c_int32(comp), C.c_int64(molecule), C.c_int64(pool_offset),
                                              C.c_double(uniform), sc, C.c_double(stored_r), C.c_int32(excl_comp), C.c_int64(excl_mol),
                                              pp, C.byref(r), C.byref(used)))
        return self._cbmc_dict(r, used.value)

    def cbmc_chain(self, cbmc_type, comp, molecule, pool_offset, uniform, excl_comp=-1, excl_mol=-1):
        r = GbCbmcResult(); used = C.c_int32()
        self._chk(self.lib.gb_cbmc_chain(self.h, C.c_int32(cbmc_type), C.c_int32(comp), C.c_int64(molecule), C.c_int64(pool_offset),
                                         C.c_double(uniform), C.c_int32(excl_comp), C.c_int64(excl_mol), C.byref(r), C.byref(used)))
        return self._cbmc_dict(r, used.value)

    def cbmc_grown_positions(self, comp):
        pos = np.zeros((int(self.system.molsize[comp]), 3))
        self._chk(self.lib.gb_cbmc_grown_positions(self.h, C.c_int32(comp), _p(pos, f64p)))
        return pos

    def reinsertion_store(self, comp):
        self._chk(self.lib.gb_reinsertion_store(self.h, C.c_int32(comp)))

    def single_body_propose(self, move_type, comp, molecule, max_change, pool_offset, want_pos=True):
        mc = (C.c_double * 3)(*max_change)
        pos = np.zeros((int(self.system.molsize[comp]), 3)) if want_pos else None
        self._chk(self.lib.gb_single_body_propose(self.h, C.c_int32(move_type), C.c_int32(comp), C.c_int64(molecule), mc, C.c_int64(pool_offset), _p(pos, f64p)))
        return pos

    def single_body_delta(self, comp, do_new=True, do_old=True):
        m = GbMoveEnergy(); ov = C.c_int32()
        self._chk(self.lib.gb_single_body_delta(self.h, C.c_int32(comp), C.c_int32(int(do_new)), C.c_int32(int(do_old)), C.byref(m), C.byref(ov)))
        return m.as_dict(), int(ov.value)

    def single_body_delta_explicit(self, comp, molid, old: TrialAtoms, new: TrialAtoms, do_new=True, do_old=True):
        m = GbMoveEnergy(); ov = C.c_int32()
        ref = new if new is not None else old
        ty = np.ascontiguousarray(ref.type.astype(np.uint64))
        self._chk(self.lib.gb_single_body_delta_explicit(self.h, C.c_int32(comp), C.c_int64(molid), C.c_int32(ref.n),
                                                         _p(old.pos, f64p) if old is not None else None, _p(new.pos, f64p) if new is not None else None,
                                                         _p(ref.scale, f64p), _p(ref.charge, f64p), _p(ref.scale_coul, f64p), _p(ty, u64p),
                                                         C.c_int32(int(do_new)), C.c_int32(int(do_old)), C.byref(m), C.byref(ov)))
        return m.as_dict(), int(ov.value)

    def ewald_delta(self, comp, move_type, location=0, scale=(1.0, 1.0)):
        sc = (C.c_double * 2)(*scale); out = np.zeros(2)
        self._chk(self.lib.gb_ewald_delta(self.h, C.c_int32(comp), C.c_int32(move_type), C.c_int64(location), sc, _p(out, f64p)))
        return out

    def ewald_delta_identity_swap(self, old_comp, new_comp, update_location):
        out = np.zeros(2)
        self._chk(self.lib.gb_ewald_delta_identity_swap(self.h, C.c_int32(old_comp), C.c_int32(new_comp), C.c_int64(update_location), _p(out, f64p)))
        return out

    def accept_translation(self, comp):
        self._chk(self.lib.gb_accept_translation(self.h, C.c_int32(comp)))

    def accept_insertion(self, comp):
        self._chk(self.lib.gb_accept_insertion(self.h, C.c_int32(comp)))

    def accept_deletion(self, comp, molecule):
        self._chk(self.lib.gb_accept_deletion(self.h, C.c_int32(comp), C.c_int64(molecule)))

    def accept_reinsertion(self, comp, molecule):
        self._chk(self.lib.gb_accept_reinsertion(self.h, C.c_int32(comp), C.c_int64(molecule)))

    def lambda_change_delta(self, comp, molecule, new_scale):
        sc = (C.c_double * 2)(*new_scale); m = GbMoveEnergy(); ov = C.c_int32()
        self._chk(self.lib.gb_lambda_change_delta(self.h, C.c_int32(comp), C.c_int64(molecule), sc, C.byref(m), C.byref(ov)))
        return m.as_dict(), int(ov.value)

    def ewald_delta_lambda_change(self, comp, old_scale, new_scale, use_temp_vector=False):
        so = (C.c_double * 2)(*old_scale); sn = (C.c_double * 2)(*new_scale); out = np.zeros(2)
        self._chk(self.lib.gb_ewald_delta_lambda_change(self.h, C.c_int32(comp), so, sn, C.c_int32(int(use_temp_vector)), _p(out, f64p)))
        return out

    def accept_lambda_change(self, comp, molecule, new_scale):
        sc = (C.c_double * 2)(*new_scale)
        self._chk(self.lib.gb_accept_lambda_change(self.h, C.c_int32(comp), C.c_int64(molecule), sc))

    def cbcf_set_scale(self, comp, molecule, scale):
        sc = (C.c_double * 2)(*scale)
        self._chk(self.lib.gb_cbcf_set_scale(self.h, C.c_int32(comp), C.c_int64(molecule), sc))

    def cbcf_deletion_stage(self, comp, molecule, revert=False):
        self._chk(self.lib.gb_cbcf_deletion_stage(self.h, C.c_int32(comp), C.c_int64(molecule), C.c_int32(int(revert))))

    def accept_identity_swap(self, old_comp, old_molecule, new_comp):
        self._chk(self.lib.gb_accept_identity_swap(self.h, C.c_int32(old_comp), C.c_int64(old_molecule), C.c_int32(new_comp)))

    def append_molecule(self, comp, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        self._chk(self.lib.gb_append_molecule(self.h, C.c_int32(comp), _p(pos, f64p), None, None, None, None))

    def number_of_molecules(self, comp):
        n = C.c_int64(); self._chk(self.lib.gb_number_of_molecules(self.h, C.c_int32(comp), C.byref(n))); return n.value

    # ------------------------------------------------------------ batched Widom
    def widom_batch(self, comp, rnd, uni, fb_index=None, or_index=None, n_blocks=5, want_outputs=True, shard=(0, 0), d_sums=None, resume=False):
        """rnd: (n_pool, 3) double3 pool (host), or None = the pool upload_random_pool left on the device; uni: (n, 2).  Packed layout
        unless indices are given.  resume: start from the first-bead energies widom_first_bead_success(None, ...) kept.
        Returns (out8 (n,8) or None, stage (n,) or None, sums (n_blocks, 12))."""
        uni = np.ascontiguousarray(uni, dtype=np.float64).reshape(-1, 2)
        if rnd is None:
            n = uni.shape[0]
            fb = np.ascontiguousarray(fb_index, dtype=np.int64) if fb_index is not None else None
            orr = np.ascontiguousarray(or_index, dtype=np.int64) if or_index is not None else None
            inp = GbWidomInputs(None, 0, fb.ctypes.data if fb is not None else None, orr.ctypes.data if orr is not None else None,
                                uni.ctypes.data, 0, n_blocks, int(shard[0]), int(shard[1]), d_sums, int(bool(resume)), 0)
            out8 = np.zeros((n, 8)) if want_outputs else None
            stage = np.zeros(n, dtype=np.int32) if want_outputs else None
            sums = np.zeros((n_blocks, 12))
            self._chk(self.lib.gb_widom_batch(self.h, C.c_int32(comp), C.c_int64(n), C.byref(inp), _p(out8, f64p), _p(stage, i32p),
                                              C.c_int32(0), _p(sums, f64p)))
            return out8, stage, sums
        rnd = np.ascontiguousarray(rnd, dtype=np.float64).reshape(-1, 3)
        n = uni.shape[0]
        fb = np.ascontiguousarray(fb_index, dtype=np.int64) if fb_index is not None else None
        orr = np.ascontiguousarray(or_index, dtype=np.int64) if or_index is not None else None
        inp = GbWidomInputs(rnd.ctypes.data, rnd.shape[0], fb.ctypes.data if fb is not None else None,
                            orr.ctypes.data if orr is not None else None, uni.ctypes.data, 0, n_blocks, int(shard[0]), int(shard[1]), d_sums)
        out8 = np.zeros((n, 8)) if want_outputs else None
        stage = np.zeros(n, dtype=np.int32) if want_outputs else None
        sums = np.zeros((n_blocks, 12))
        self._chk(self.lib.gb_widom_batch(self.h, C.c_int32(comp), C.c_int64(n), C.byref(inp), _p(out8, f64p), _p(stage, i32p),
                                          C.c_int32(0), _p(sums, f64p)))
        return out8, stage, sums

    def widom_first_bead_success(self, comp, rnd, fb_index):
        """first-bead classification of the pool blocks starting at fb_index (1 success, 0 failed with a survivor, 2 no survivor);
        rnd None = the pool upload_random_pool left on the device"""
        fb = np.ascontiguousarray(fb_index, dtype=np.int64); code = np.zeros(fb.shape[0], dtype=np.int32)
        if rnd is not None:
            rnd = np.ascontiguousarray(rnd, dtype=np.float64).reshape(-1, 3)
        self._chk(self.lib.gb_widom_first_bead_success(self.h, C.c_int32(comp), C.c_int64(fb.shape[0]), _p(rnd, f64p) if rnd is not None else None,
                                                       C.c_int64(rnd.shape[0] if rnd is not None else 0), _p(fb, i64p), _p(code, i32p)))
        return code

    def widom_batch_device(self, comp, n, d_pool, n_pool, d_uni, n_blocks=5, d_out8=None, shard=(0, 0), d_sums=None, want_host_sums=True):
        """device-resident inputs (raw device pointers as ints, e.g. torch tensor.data_ptr()); d_sums: device pointer of n_blocks*12
        doubles the block sums are ADDED to on the engine's stream (gb_widom_inputs.sums_device)"""
        inp = GbWidomInputs(d_pool, n_pool, None, None, d_uni, 1, n_blocks, int(shard[0]), int(shard[1]), d_sums)
        sums = np.zeros((n_blocks, 12)) if want_host_sums else None
        self._chk(self.lib.gb_widom_batch(self.h, C.c_int32(comp), C.c_int64(n), C.byref(inp), C.c_void_p(d_out8) if d_out8 else None,
                                          None, C.c_int32(1), _p(sums, f64p)))
        return sums

    # ------------------------------------------------------------ fused moves (one kernel, one host round trip per move)
    def move_insertion(self, comp, pool_offset, uniforms, scale=(1.0, 1.0)):
        u = (C.c_double * 2)(*uniforms); sc = (C.c_double * 2)(*scale); r = GbMoveResult()
        self._chk(self.lib.gb_move_insertion(self.h, C.c_int32(comp), C.c_int64(pool_offset), u, sc, C.byref(r))); return r.as_dict()

    def move_deletion(self, comp, molecule, pool_offset, scale=(1.0, 1.0)):
        sc = (C.c_double * 2)(*scale); r = GbMoveResult()
        self._chk(self.lib.gb_move_deletion(self.h, C.c_int32(comp), C.c_int64(molecule), C.c_int64(pool_offset), sc, C.byref(r))); return r.as_dict()

    def move_reinsertion(self, comp, molecule, pool_offset, uniforms):
        u = (C.c_double * 2)(*uniforms); r = GbMoveResult()
        self._chk(self.lib.gb_move_reinsertion(self.h, C.c_int32(comp), C.c_int64(molecule), C.c_int64(pool_offset), u, C.byref(r))); return r.as_dict()

    def move_identity_swap(self, old_comp, old_molecule, new_comp, pool_offset, uniform):
        r = GbMoveResult()
        self._chk(self.lib.gb_move_identity_swap(self.h, C.c_int32(old_comp), C.c_int64(old_molecule), C.c_int32(new_comp), C.c_int64(pool_offset),
                                                 C.c_double(uniform), C.byref(r)))
        return r.as_dict()

    def move_single_body(self, move_type, comp, molecule, max_change, pool_offset):
        mc = (C.c_double * 3)(*max_change); r = GbMoveResult()
        self._chk(self.lib.gb_move_single_body(self.h, C.c_int32(move_type), C.c_int32(comp), C.c_int64(molecule), mc, C.c_int64(pool_offset), C.byref(r)))
        return r.as_dict()

    def stream(self):
        """cudaStream_t of the engine as an int (wrap with torch.cuda.ExternalStream to record events on it)"""
        return int(self.lib.gb_stream(self.h) or 0)

    # ------------------------------------------------------------ instrumentation
    def launch_count(self, reset=False):
        n = C.c_int64(); self._chk(self.lib.gb_launch_count(self.h, C.byref(n), C.c_int32(int(reset)))); return n.value

    def timing_enable(self, on=True):
        self._chk(self.lib.gb_timing_enable(self.h, C.c_int32(int(on))))

    def timing_read(self, family, reset=False):
        ms = C.c_double(); n = C.c_int64()
        self._chk(self.lib.gb_timing_read(self.h, C.c_int32(family), C.byref(ms), C.byref(n), C.c_int32(int(reset))))
        return ms.value, n.value

    def move_server(self, on=None):
        """enable / disable the resident move server (None: leave as is); returns (server launches, moves it executed)"""
        a = C.c_int64(); b = C.c_int64()
        self._chk(self.lib.gb_move_server(self.h, C.c_int32(-1 if on is None else int(bool(on))), C.byref(a), C.byref(b)))
        return a.value, b.value

    def measure_fp64_peak(self):
        t = C.c_double(); self._chk(self.lib.gb_measure_fp64_peak(self.h, C.byref(t))); return t.value

    def synchronize(self):
        self._chk(self.lib.gb_synchronize(self.h))
