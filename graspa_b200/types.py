"""Plain host-side containers for the engine's inputs.

They mirror the reference's host structs field for field (paths relative to
/root/reference/src_clean):

* ``Box``        <- ``Boxsize``     data_struct.h:865-886
* ``ForceField`` <- ``ForceField``  data_struct.h:838-855 (+ ``Tail`` :720-724)
* ``System``     <- ``Atoms[]``     data_struct.h:788-799, all components concatenated
* ``MoveEnergy`` field order        data_struct.h:416-431

No arithmetic happens here beyond what the reference does on the host at set-up
time (cell inverse/determinant, maths.cuh:28-56).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

COULOMB_PREFACTOR = 138935.483496  # read_data.cpp:611

# data_struct.h:20
TRANSLATION, ROTATION, SINGLE_INSERTION, SINGLE_DELETION, SPECIAL_ROTATION, INSERTION, DELETION, \
    REINSERTION, CBCF_LAMBDACHANGE, CBCF_INSERTION, CBCF_DELETION, IDENTITY_SWAP, WIDOM = range(13)
# data_struct.h:22
CBMC_INSERTION, CBMC_DELETION, REINSERTION_INSERTION, REINSERTION_RETRACE, IDENTITY_SWAP_NEW, IDENTITY_SWAP_OLD = range(6)


def beta_from_temperature(T: float) -> float:
    """fxn_main.h:117 with the ``Units`` constants of data_struct.h:58-68."""
    kB, mass_unit, length_unit, time_unit = 1.380649e-23, 1.6605402e-27, 1e-10, 1e-12
    return 1.0 / (kB / (mass_unit * pow(length_unit, 2) / pow(time_unit, 2)) * T)


def _f64(a, shape=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    if shape is not None:
        a = a.reshape(shape)
    return a


def inverse_cell(cell):
    """maths.cuh:38-56 (cofactor inverse of the row-vector cell) and :28-36 (determinant)."""
    x = _f64(cell, (9,))
    m11, m12, m13, m21, m22, m23, m31, m32, m33 = x
    det = +m11 * (m22 * m33 - m23 * m32) - m12 * (m21 * m33 - m23 * m31) + m13 * (m21 * m32 - m22 * m31)
    r = np.empty(9)
    r[0] = +(m22 * m33 - m32 * m23) / det
    r[3] = -(m21 * m33 - m31 * m23) / det
    r[6] = +(m21 * m32 - m31 * m22) / det
    r[1] = -(m12 * m33 - m32 * m13) / det
    r[4] = +(m11 * m33 - m31 * m13) / det
    r[7] = -(m11 * m32 - m31 * m12) / det
    r[2] = +(m12 * m23 - m22 * m13) / det
    r[5] = -(m11 * m23 - m21 * m13) / det
    r[8] = +(m11 * m22 - m21 * m12) / det
    return r, float(det)


@dataclass
class Box:
    cell: np.ndarray                      # 9, rows = lattice vectors (lower triangular)
    alpha: float = 0.0
    kmax: tuple = (0, 0, 0)
    recip_cutoff: float = 0.0
    prefactor: float = COULOMB_PREFACTOR
    use_lammps_ewald: bool = False
    inv: np.ndarray = field(default=None)
    volume: float = 0.0
    cubic: bool = False

    def __post_init__(self):
        self.cell = _f64(self.cell, (9,))
        if self.inv is None:
            self.inv, self.volume = inverse_cell(self.cell)
        self.inv = _f64(self.inv, (9,))
        # read_data.cpp:2024-2025
        self.cubic = not ((abs(self.cell[3]) + abs(self.cell[6]) + abs(self.cell[7])) > 1e-10)

    @property
    def nvec(self) -> int:
        return (self.kmax[0] + 1) * (2 * self.kmax[1] + 1) * (2 * self.kmax[2] + 1)


@dataclass
class ForceField:
    epsilon: np.ndarray                   # ntypes*ntypes, already / 1.20272430057 (read_data.cpp:1200)
    sigma: np.ndarray
    shift: np.ndarray
    cutoff_vdw: float
    cutoff_coul: float
    overlap: float = 1e5
    no_charges: bool = False
    vdw_real_bias: bool = True            # data_struct.h:1370 / SURVEY section 5 config quirk
    use1264: bool = False
    z: Optional[np.ndarray] = None
    c10: Optional[np.ndarray] = None
    use_tail: Optional[np.ndarray] = None  # int32 ntypes*ntypes
    tail_energy: Optional[np.ndarray] = None

    def __post_init__(self):
        self.epsilon = _f64(self.epsilon).ravel()
        n2 = self.epsilon.size
        self.ntypes = int(round(np.sqrt(n2)))
        assert self.ntypes * self.ntypes == n2
        self.sigma = _f64(self.sigma).ravel()
        self.shift = _f64(self.shift).ravel()
        self.z = _f64(self.z if self.z is not None else np.zeros(n2)).ravel()
        self.c10 = _f64(self.c10 if self.c10 is not None else np.zeros(n2)).ravel()
        self.use_tail = np.ascontiguousarray(
            self.use_tail if self.use_tail is not None else np.zeros(n2), dtype=np.int32).ravel()
        self.tail_energy = _f64(self.tail_energy if self.tail_energy is not None else np.zeros(n2)).ravel()

    @property
    def cutoff_vdw_sq(self):
        return self.cutoff_vdw * self.cutoff_vdw

    @property
    def cutoff_coul_sq(self):
        return self.cutoff_coul * self.cutoff_coul

    @property
    def has_tail(self):
        return bool(self.use_tail.any())


@dataclass
class System:
    """All components concatenated.  Component c owns slots [offset[c], offset[c]+alloc[c]); the first
    natoms[c] are live.  Slot 0 of an adsorbate component always holds a molecule (the .def template
    when the component is empty) because the reference grows chains from ``d_a[c].pos[1+a]-pos[0]``
    (mc_widom.h:256, read_data.cpp:2122-2147)."""
    nhost: int
    natoms: np.ndarray                    # int64 per component (live atoms)
    molsize: np.ndarray                   # int64 per component
    pos: np.ndarray                       # (sum(alloc),3)
    charge: np.ndarray
    type: np.ndarray                      # int64
    molid: np.ndarray                     # int64
    scale: Optional[np.ndarray] = None
    scale_coul: Optional[np.ndarray] = None
    alloc: Optional[np.ndarray] = None    # int64 per component (slots), default max(natoms, molsize)

    def __post_init__(self):
        self.natoms = np.ascontiguousarray(self.natoms, dtype=np.int64)
        self.molsize = np.ascontiguousarray(self.molsize, dtype=np.int64)
        self.ncomp = int(self.natoms.size)
        if self.alloc is None:
            self.alloc = np.maximum(self.natoms, self.molsize)
        self.alloc = np.ascontiguousarray(self.alloc, dtype=np.int64)
        n = int(self.alloc.sum())
        self.pos = _f64(self.pos, (n, 3))
        self.charge = _f64(self.charge, (n,))
        self.type = np.ascontiguousarray(self.type, dtype=np.int64).reshape(n)
        self.molid = np.ascontiguousarray(self.molid, dtype=np.int64).reshape(n)
        self.scale = _f64(self.scale if self.scale is not None else np.ones(n), (n,))
        self.scale_coul = _f64(self.scale_coul if self.scale_coul is not None else np.ones(n), (n,))

    @property
    def offsets(self):
        return np.concatenate([[0], np.cumsum(self.alloc)]).astype(np.int64)

    @property
    def nslots(self):
        return int(self.alloc.sum())

    def component(self, c):
        """slice of the LIVE atoms of component c"""
        o = self.offsets
        return slice(int(o[c]), int(o[c] + self.natoms[c]))

    def slots(self, c):
        o = self.offsets
        return slice(int(o[c]), int(o[c + 1]))

    def live_mask(self):
        m = np.zeros(self.nslots, dtype=bool)
        for c in range(self.ncomp):
            m[self.component(c)] = True
        return m

    def compact(self) -> "System":
        """live atoms only (alloc == natoms); what the reference's host totals loop over"""
        m = self.live_mask()
        return System(self.nhost, self.natoms.copy(), self.molsize.copy(), self.pos[m], self.charge[m],
                      self.type[m], self.molid[m], self.scale[m], self.scale_coul[m], alloc=self.natoms.copy())


@dataclass
class TrialAtoms:
    """A set of trial / moved atoms (``Sims.New`` / ``Sims.Old``)."""
    pos: np.ndarray
    charge: np.ndarray
    type: np.ndarray
    scale: Optional[np.ndarray] = None
    scale_coul: Optional[np.ndarray] = None

    def __post_init__(self):
        self.pos = _f64(self.pos).reshape(-1, 3)
        n = self.pos.shape[0]
        self.n = n
        self.charge = _f64(self.charge, (n,))
        self.type = np.ascontiguousarray(self.type, dtype=np.int64).reshape(n)
        self.scale = _f64(self.scale if self.scale is not None else np.ones(n), (n,))
        self.scale_coul = _f64(self.scale_coul if self.scale_coul is not None else np.ones(n), (n,))


def species_counts(system: System, comp: int, ntypes: int) -> np.ndarray:
    """Atoms of each pseudo-atom type in one molecule of ``comp``
    (``NumberOfPseudoAtomsForSpecies``, TailCorrection_Energy_Functions.h:23-34)."""
    o = int(system.offsets[comp])
    ms = int(system.molsize[comp])
    t = system.type[o:o + ms]
    return np.bincount(t, minlength=ntypes).astype(np.int32)


def pseudo_atom_counts(system: System, ntypes: int) -> np.ndarray:
    """``NumberOfPseudoAtoms`` (live atoms of each type in the whole box)."""
    return np.bincount(system.type[system.live_mask()], minlength=ntypes).astype(np.int64)
