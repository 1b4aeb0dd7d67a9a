"""The oracle's Monte Carlo logic against values written by the REFERENCE PROGRAM itself while it ran on a B200
(oracle/build_ref.sh dump: the reference's CUDA program instrumented in a scratch copy; scripts/make_ref_dump.sh ran it;
tests/golden/make_ref_dump.py made the fixtures).  This pins, with reference-derived numbers, the functions VERDICT r1 found
unpinned: orc_trial_positions / _orientations / orc_rotate_quaternions, orc_select_trial / orc_cbmc_finish, orc_blocked_pocket,
and the whole Widom insertion (so the widom_out column of the config fixtures is no longer oracle-only).  CPU only."""
import os

import numpy as np
import pytest

from graspa_b200.types import Box, TrialAtoms, CBMC_INSERTION
from tests.conftest import load_config, GOLDEN


def _dump(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def test_trial_positions_and_orientations_match_the_reference_kernels(oracle):
    """get_random_trial_position (mc_widom.h:122-213), get_random_trial_orientation + Rotate_Quaternions (:215-303,
    mc_utilities.h:423-457): positions the reference's kernels produced from the same pool randoms, 150 insertions"""
    box, ff, s, z = load_config("A")
    d = _dump("ref_dump_widom_A.npz")
    comp = int(z["comp"]); ms = int(s.molsize[comp])
    for k in range(len(d["fb_off"])):
        tr = oracle.trial_positions(box, s, CBMC_INSERTION, comp, 0, 10, d["fb_rnd"][k])
        assert np.max(np.abs(tr.pos - d["fb_pos"][k])) < 1e-11
        t2 = oracle.trial_orientations(s, CBMC_INSERTION, comp, 1, ms - 1, 10, d["ch_rnd"][k], d["ch_fb"][k])
        assert np.max(np.abs(t2.pos - d["ch_pos"][k])) < 1e-11
        # the first bead the chain grows from is the selected trial
        sel = int(d["fbe"][k][int(d["fbs"][k][1]), 0])
        assert np.array_equal(d["ch_fb"][k], d["fb_pos"][k][sel])


def test_trial_energies_match_the_reference_cuda_kernel(oracle):
    """Calculate_Multiple_Trial_Energy_VDWReal + Host_sum_Widom_HGGG_SEPARATE as the reference ran them on the GPU: per-trial HG
    VDW / real energies and the set of surviving trials, first bead and chain, 60 insertions"""
    box, ff, s, z = load_config("A")
    d = _dump("ref_dump_widom_A.npz")
    comp = int(z["comp"]); ms = int(s.molsize[comp]); new_molid = int(s.natoms[comp]) // ms
    o = int(s.offsets[comp])
    for k in range(60):
        tr = TrialAtoms(d["fb_pos"][k], np.full(10, s.charge[o]), np.full(10, s.type[o]))
        e, f, _ = oracle.trial_energies(box, ff, s, 10, 1, tr, comp, new_molid)
        n = int(d["fbe_n"][k]); rows = d["fbe"][k][:n]
        assert sorted(np.flatnonzero(f == 0).tolist()) == rows[:, 0].astype(int).tolist()
        for r in rows:
            t = int(r[0])
            assert np.max(np.abs(e[t] - r[2:6])) <= 1e-10 * max(1.0, float(np.abs(r[2:6]).sum())), (k, t, e[t], r)
            tot = e[t].sum()
            assert abs(-float(z["beta"]) * tot - r[1]) <= 1e-10 * max(1.0, abs(r[1]))
        t2 = TrialAtoms(d["ch_pos"][k], np.tile(s.charge[o + 1:o + ms], 10), np.tile(s.type[o + 1:o + ms], 10))
        e2, f2, _ = oracle.trial_energies(box, ff, s, 10, ms - 1, t2, comp, new_molid)
        n2 = int(d["che_n"][k]); rows2 = d["che"][k][:n2]
        assert sorted(np.flatnonzero(f2 == 0).tolist()) == rows2[:, 0].astype(int).tolist()
        for r in rows2:
            t = int(r[0])
            assert np.max(np.abs(e2[t] - r[2:6])) <= 1e-10 * max(1.0, float(np.abs(r[2:6]).sum()))


def test_boltzmann_selection_matches_the_reference(oracle):
    """SelectTrialPosition + the Rosenbluth sums of CBMC_FirstBead_Finish / Widom_Move_Chain_PARTIAL (mc_widom.h:14-39, 305-383,
    568-611) on 817 selections the reference made in CO2-MFI, CO2_NaX_Zeolite and BlockPocket runs (1 to 10 survivors) and 300 in
    the Henry run: same selected trial from the same log Boltzmann factors and uniform, same sum"""
    from oracle.oracle import lib, _p, f64p
    import ctypes as C
    L = lib()
    L.orc_select_trial.restype = C.c_int
    recs = []
    d = _dump("ref_dump_select.npz")
    for lb, m in zip(d["logs"], d["meta"]):
        recs.append((lb[:int(m[0])], m[1], int(m[2]), int(m[3]), m[4]))
    w = _dump("ref_dump_widom_A.npz")
    for k in range(len(w["fb_off"])):
        for e, n, srec in ((w["fbe"][k], int(w["fbe_n"][k]), w["fbs"][k]), (w["che"][k], int(w["che_n"][k]), w["chs"][k])):
            recs.append((e[:n, 1], srec[3], int(srec[0]), int(srec[1]), srec[2]))
    assert len(recs) > 1000
    few = 0
    for lb, u, good, sel, rsum in recs:
        lb = np.ascontiguousarray(lb, dtype=np.float64)
        few += len(lb) <= 3
        if u >= 0.0:                                        # a uniform was drawn (insertion-type stage with survivors)
            got = L.orc_select_trial(_p(lb, f64p), C.c_int(len(lb)), C.c_double(float(u)))
            assert got == sel, (lb, u, got, sel)
        total = float(np.sum(np.exp(lb)))                   # exp then accumulate in trial order (:336-337)
        acc = 0.0
        for x in lb:
            acc += float(np.exp(x))
        assert abs(acc - rsum) <= 1e-12 * max(abs(rsum), 1e-300), (acc, rsum)
        assert good == (1 if rsum >= 1e-150 else 0)
        rosen = lb.copy(); sel_o = C.c_int(0); R = C.c_double(0.0); st = C.c_double(0.0)
        ok = L.orc_cbmc_finish(C.c_int(CBMC_INSERTION), C.c_int(0), _p(rosen, f64p), C.c_int(len(lb)), C.c_int(1), C.c_double(float(max(u, 0.0))),
                               C.c_double(0.0), C.byref(st), C.byref(sel_o), C.byref(R))
        assert bool(ok) == bool(good)
        if good:
            assert abs(R.value - rsum) <= 1e-12 * abs(rsum) and (u < 0.0 or sel_o.value == sel)
    assert few > 100


def test_whole_insertions_match_the_reference_program(oracle):
    """Insertion_Body (mc_swap_utilities.h:3-133) end to end: final Rosenbluth weight within 1e-9 relative (BASELINE.json) and the
    energy terms incl. the Fourier and exclusion parts, for the first 150 Widom insertions of Examples/Henrys_coefficient, seed 0"""
    box, ff, s, z = load_config("A")
    d = _dump("ref_dump_widom_A.npz")
    comp = int(z["comp"])
    ws = oracle.WidomSetup(box, ff, s, comp, float(z["beta"]), 10, 10, z["sf_ads"], z["sf_fw"])
    n = len(d["fb_off"])
    assert d["has_ins"].all()
    rnd = np.concatenate([d["fb_rnd"], d["ch_rnd"]], axis=1)
    uni = np.stack([d["fbs"][:, 3], d["chs"][:, 3]], axis=1)
    out, stage, _ = oracle.widom_batch(ws, rnd, uni)
    assert (stage == 0).all()
    ref = d["ins"]
    assert np.max(np.abs(out[:, 0] - ref[:, 0]) / ref[:, 0]) < 1e-9
    # out8 = {W, HGVDW, HGReal, GGVDW, GGReal, GGEwaldE, HGEwaldE, TailE}: the dump carries the same order
    esc = np.abs(ref[:, 1:]).sum(axis=1, keepdims=True) + 1e-3
    assert np.max(np.abs(out[:, 1:] - ref[:, 1:]) / esc) < 1e-9
    # the pool offsets advance by 10 + 10 per insertion: the RNG bookkeeping of SURVEY 9.1
    assert (np.diff(d["fb_off"]) == 20).all() and (d["ch_off"] - d["fb_off"] == 10).all()


@pytest.mark.parametrize("deck", ["CO2_NaX_Zeolite", "BlockPocket"])
def test_blocked_pocket_matches_the_reference(oracle, deck):
    """BlockedPocket (read_data.cpp:3466-3640): the reference's verdicts for positions it tested during the runs"""
    p = _dump("ref_dump_pockets.npz")
    box = Box(p[f"{deck}_cell"])
    calls = p[f"{deck}_calls"]
    assert (calls[:, 3] > 0.5).sum() > 50 and (calls[:, 3] < 0.5).sum() > 50
    for c in calls:
        got = oracle.blocked_pocket(box, p[f"{deck}_centers"], p[f"{deck}_radii"], c[:3], invert=bool(p[f"{deck}_invert"]))
        assert got == (c[3] > 0.5), c
