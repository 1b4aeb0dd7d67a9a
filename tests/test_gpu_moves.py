"""Single-move path of the CUDA engine (CBMC stages, translation/rotation, Ewald deltas by move type, state commits)
through the C ABI against the oracle, plus the reference's own self-check: running sum of accepted move deltas
versus total energies recomputed from scratch (ENERGY DRIFT, fxn_main.h:467-468, test_examples.py:59-61)."""
import numpy as np
import pytest

from graspa_b200.types import (TrialAtoms, CBMC_INSERTION, CBMC_DELETION, REINSERTION_INSERTION, REINSERTION_RETRACE,
                               TRANSLATION, ROTATION, INSERTION, DELETION, REINSERTION, CBCF_INSERTION, CBCF_DELETION)
from graspa_b200.types import pseudo_atom_counts
from tests.conftest import load_config

pytestmark = pytest.mark.gpu
ETOL = 1e-10


def _grow(s, comp, nslots):
    """same system with more allocated slots for one component (AdsorbateAllocateSpace)"""
    from graspa_b200.types import System
    alloc = s.alloc.copy(); alloc[comp] = nslots
    off_o = s.offsets; off_n = np.concatenate([[0], np.cumsum(alloc)])
    n = int(alloc.sum())
    pos = np.zeros((n, 3)); q = np.zeros(n); ty = np.zeros(n, dtype=np.int64); mo = np.zeros(n, dtype=np.int64)
    for c in range(s.ncomp):
        k = int(s.alloc[c])
        pos[off_n[c]:off_n[c] + k] = s.pos[off_o[c]:off_o[c] + k]; q[off_n[c]:off_n[c] + k] = s.charge[off_o[c]:off_o[c] + k]
        ty[off_n[c]:off_n[c] + k] = s.type[off_o[c]:off_o[c] + k]; mo[off_n[c]:off_n[c] + k] = s.molid[off_o[c]:off_o[c] + k]
    return System(s.nhost, s.natoms.copy(), s.molsize.copy(), pos, q, ty, mo, alloc=alloc)


def _setup(gpu_engine_factory, name="B", grow=None):
    box, ff, s, z = load_config(name)
    if grow:
        s = _grow(s, int(z["comp"]), grow)
    eng = gpu_engine_factory(box, ff, s, float(z["beta"]), int(z["ntrials"]), int(z["norient"]))
    if "sf_ads" in z:
        eng.upload_structure_factors(z["sf_ads"], z["sf_fw"])
        eng.set_exclusion_constants(int(z["comp"]), float(z["excl"][0]), float(z["excl"][1]))
    return box, ff, s, z, eng


def _close(a, b, scale=None, tol=ETOL):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    sc = scale if scale is not None else max(1e-3, float(np.abs(b).sum()))
    return np.max(np.abs(a - b)) / sc < tol


def test_cbmc_insertion_stages_vs_oracle(gpu_engine_factory, oracle):
    box, ff, s, z, eng = _setup(gpu_engine_factory)
    comp = 1; ms = 3; beta = float(z["beta"]); new_molid = int(s.natoms[comp]) // ms
    rng = np.random.default_rng(21)
    pool = rng.random((512, 3)); eng.upload_random_pool(pool)
    for rep in range(6):
        off = 40 * rep; u1, u2 = rng.random(2)
        fb = eng.cbmc_first_bead(CBMC_INSERTION, comp, 0, off, u1)
        tr = oracle.trial_positions(box, s, CBMC_INSERTION, comp, 0, 10, pool[off:off + 10])
        e, f, _ = oracle.trial_energies(box, ff, s, 10, 1, tr, comp, new_molid)
        ref = oracle.cbmc_finish(CBMC_INSERTION, False, e, f, beta, 10, u1)
        assert fb["success"] == ref["success"] and fb["n_survivors"] == ref["nsurv"]
        if not ref["success"]:
            continue
        assert fb["selected"] == ref["selected"] and fb["uniform_used"] == 1
        assert abs(fb["rosenbluth"] - ref["rosenbluth"]) <= 1e-9 * ref["rosenbluth"]
        assert _close(fb["energy"], e[ref["selected"]])
        assert np.allclose(fb["selected_pos"], tr.pos[ref["selected"]], rtol=0, atol=1e-12)
        ch = eng.cbmc_chain(CBMC_INSERTION, comp, 0, off + 10, u2)
        t2 = oracle.trial_orientations(s, CBMC_INSERTION, comp, 1, ms - 1, 10, pool[off + 10:off + 20], tr.pos[ref["selected"]])
        e2, f2, _ = oracle.trial_energies(box, ff, s, 10, ms - 1, t2, comp, new_molid)
        ref2 = oracle.cbmc_finish(CBMC_INSERTION, True, e2, f2, beta, 10, u2)
        assert ch["success"] == ref2["success"]
        if not ref2["success"]:
            continue
        assert ch["selected"] == ref2["selected"]
        assert abs(ch["rosenbluth"] - ref2["rosenbluth"]) <= 1e-9 * ref2["rosenbluth"]
        assert _close(ch["energy"], e2[ref2["selected"]])
        grown = eng.cbmc_grown_positions(comp)
        so = ref2["selected"]
        expect = np.concatenate([tr.pos[ref["selected"]][None, :], t2.pos[so * 2:so * 2 + 2]])
        assert np.allclose(grown, expect, rtol=0, atol=1e-11)
        # Ewald delta of the grown molecule (GPU_EwaldDifference_General, INSERTION)
        o = int(s.offsets[comp]); q = s.charge[o:o + ms]
        got = eng.ewald_delta(comp, INSERTION, location=ch["selected"])
        ew, _, _ = oracle.ewald_delta(box, expect, q, np.ones(ms), 0, ms, z["sf_ads"], z["sf_fw"])
        ew[0] -= float(z["excl"][0]) + float(z["excl"][1])
        assert _close(got, ew, scale=max(1.0, float(np.abs(ew).max())))
    eng.close()


def test_cbmc_deletion_and_reinsertion_stages_vs_oracle(gpu_engine_factory, oracle):
    box, ff, s, z, eng = _setup(gpu_engine_factory)
    comp = 1; ms = 3; beta = float(z["beta"])
    rng = np.random.default_rng(22)
    pool = rng.random((512, 3)); eng.upload_random_pool(pool)
    o = int(s.offsets[comp])
    for mol in (0, 7, 19):
        off = 30 * mol % 400
        fb = eng.cbmc_first_bead(CBMC_DELETION, comp, mol, off, 0.5)
        tr = oracle.trial_positions(box, s, CBMC_DELETION, comp, mol * ms, 10, pool[off:off + 10])
        assert np.allclose(tr.pos[0], s.pos[o + mol * ms])
        e, f, _ = oracle.trial_energies(box, ff, s, 10, 1, tr, comp, mol)
        ref = oracle.cbmc_finish(CBMC_DELETION, False, e, f, beta, 10, 0.5)
        assert fb["success"] and fb["selected"] == 0 and fb["uniform_used"] == 0
        assert abs(fb["rosenbluth"] - ref["rosenbluth"]) <= 1e-9 * ref["rosenbluth"]
        ch = eng.cbmc_chain(CBMC_DELETION, comp, mol, off + 10, 0.5)
        t2 = oracle.trial_orientations(s, CBMC_DELETION, comp, mol * ms + 1, ms - 1, 10, pool[off + 10:off + 20], tr.pos[0])
        assert np.allclose(t2.pos[:2], s.pos[o + mol * ms + 1:o + mol * ms + 3])
        e2, f2, _ = oracle.trial_energies(box, ff, s, 10, ms - 1, t2, comp, mol)
        ref2 = oracle.cbmc_finish(CBMC_DELETION, True, e2, f2, beta, 10, 0.5)
        assert ch["selected"] == 0 and abs(ch["rosenbluth"] - ref2["rosenbluth"]) <= 1e-9 * ref2["rosenbluth"]
        assert _close(ch["energy"], e2[0])
        got = eng.ewald_delta(comp, DELETION, location=mol * ms)
        q = s.charge[o:o + ms]
        ew, _, _ = oracle.ewald_delta(box, s.pos[o + mol * ms:o + mol * ms + ms], q, np.ones(ms), ms, 0, z["sf_ads"], z["sf_fw"])
        ew[0] -= (float(z["excl"][0]) + float(z["excl"][1])) * -1.0
        assert _close(got, ew, scale=max(1.0, float(np.abs(ew).max())))
    # reinsertion: insertion leg with StoredR, store, retrace leg with 1 trial + StoredR, Ewald(REINSERTION)
    mol = 5; off = 200; u1, u2 = 0.37, 0.81
    fb = eng.cbmc_first_bead(REINSERTION_INSERTION, comp, mol, off, u1)
    tr = oracle.trial_positions(box, s, REINSERTION_INSERTION, comp, mol * ms, 10, pool[off:off + 10])
    e, f, _ = oracle.trial_energies(box, ff, s, 10, 1, tr, comp, mol)
    ref = oracle.cbmc_finish(REINSERTION_INSERTION, False, e, f, beta, 10, u1)
    assert fb["success"] == ref["success"]
    if ref["success"]:
        assert fb["selected"] == ref["selected"]
        assert abs(fb["stored_r"] - ref["stored_r"]) <= 1e-9 * max(abs(ref["stored_r"]), 1e-300)
        ch = eng.cbmc_chain(REINSERTION_INSERTION, comp, mol, off + 10, u2)
        if ch["success"]:
            new_pos = eng.cbmc_grown_positions(comp)
            eng.reinsertion_store(comp)
            rb = eng.cbmc_first_bead(REINSERTION_RETRACE, comp, mol, off + 20, 0.5, stored_r=fb["stored_r"])
            tr_r = oracle.trial_positions(box, s, REINSERTION_RETRACE, comp, mol * ms, 1, pool[off + 20:off + 21])
            er, fr, _ = oracle.trial_energies(box, ff, s, 1, 1, tr_r, comp, mol)
            ref_r = oracle.cbmc_finish(REINSERTION_RETRACE, False, er, fr, beta, 10, 0.5, stored_in=ref["stored_r"])
            assert abs(rb["rosenbluth"] - ref_r["rosenbluth"]) <= 1e-9 * ref_r["rosenbluth"]
            eng.cbmc_chain(REINSERTION_RETRACE, comp, mol, off + 21, 0.5)
            got = eng.ewald_delta(comp, REINSERTION, location=mol * ms)
            q = s.charge[o:o + ms]
            pos = np.concatenate([s.pos[o + mol * ms:o + mol * ms + ms], new_pos])
            ew, _, _ = oracle.ewald_delta(box, pos, np.concatenate([q, q]), np.ones(2 * ms), ms, ms, z["sf_ads"], z["sf_fw"])
            assert _close(got, ew, scale=max(1.0, float(np.abs(ew).max())))
    eng.close()


def _rot(p, theta, axis):
    c, s_ = np.cos(theta), np.sin(theta); w = 1.0 - c; ax, ay, az = axis
    R = np.array([[ax * ax * w + c, ax * ay * w + az * s_, ax * az * w - ay * s_],
                  [ax * ay * w - az * s_, ay * ay * w + c, ay * az * w + ax * s_],
                  [ax * az * w + ay * s_, ay * az * w - ax * s_, az * az * w + c]])
    return R @ p


def test_single_body_translation_rotation_vs_oracle(gpu_engine_factory, oracle):
    box, ff, s, z, eng = _setup(gpu_engine_factory)
    comp = 1; ms = 3; o = int(s.offsets[comp])
    rng = np.random.default_rng(23)
    pool = rng.random((64, 3)); eng.upload_random_pool(pool)
    for k, (mt, mol) in enumerate([(TRANSLATION, 3), (ROTATION, 11), (TRANSLATION, 19), (ROTATION, 0)]):
        maxc = np.array([0.8, 0.6, 0.7]) if mt == TRANSLATION else np.array([0.5, 0.4, 0.3])
        newp = eng.single_body_propose(mt, comp, mol, maxc, k)
        oldp = s.pos[o + mol * ms:o + mol * ms + ms]
        r = pool[k]
        if mt == TRANSLATION:
            expect = oldp + maxc * 2.0 * (r - 0.5)
        else:
            ang = maxc * 2.0 * (r - 0.5)
            expect = np.array([_rot(_rot(_rot(p - oldp[0], ang[0], (1, 0, 0)), ang[1], (0, 1, 0)), ang[2], (0, 0, 1)) + oldp[0] for p in oldp])
        assert np.allclose(newp, expect, rtol=0, atol=1e-11)
        d, ov = eng.single_body_delta(comp)
        q = s.charge[o:o + ms]; ty = s.type[o:o + ms]
        ref, rov = oracle.single_body_delta(box, ff, s, comp, mol, TrialAtoms(oldp, q, ty), TrialAtoms(newp, q, ty))
        got = np.array([d["HHVDW"], d["HHReal"], d["HGVDW"], d["HGReal"], d["GGVDW"], d["GGReal"]])
        assert ov == rov
        # tolerance against the magnitude of the summed terms (new and old each ~1e3-1e4 K), SURVEY section 7 "tolerance vs cancellation"
        assert np.max(np.abs(got - ref)) < ETOL * 1e4
        ew = eng.ewald_delta(comp, mt)
        ewr, _, _ = oracle.ewald_delta(box, np.concatenate([oldp, newp]), np.concatenate([q, q]), np.ones(2 * ms), ms, ms, z["sf_ads"], z["sf_fw"])
        assert _close(ew, ewr, scale=max(1.0, float(np.abs(ewr).max())))
        # explicit-atoms entry point gives the same numbers
        d2, _ = eng.single_body_delta_explicit(comp, mol, TrialAtoms(oldp, q, ty), TrialAtoms(newp, q, ty))
        assert all(abs(d2[k2] - d[k2]) <= 1e-9 * max(1.0, abs(d[k2])) for k2 in d)
    eng.close()


def _total(eng):
    v = eng.total_vdw_real(); w = eng.total_ewald(store=False)
    return (v["HHVDW"] + v["HGVDW"] + v["GGVDW"] + v["HHReal"] + v["HGReal"] + v["GGReal"] + w["GGEwaldE"] + w["HGEwaldE"] + eng.tail_total())


def test_energy_drift_over_a_short_gcmc_run(gpu_engine_factory):
    """insert / delete / translate / rotate / reinsert with Metropolis acceptance on the engine's own deltas;
    the running sum of accepted deltas must equal the from-scratch total-energy difference (drift < 1e-3 in the
    reference's criterion; here 1e-7 relative to the energy scale)"""
    box, ff, s, z, eng = _setup(gpu_engine_factory, grow=600)
    comp = 1; ms = 3; beta = float(z["beta"])
    rng = np.random.default_rng(24)
    pool = rng.random((20000, 3)); eng.upload_random_pool(pool)
    E0 = _total(eng)
    run = 0.0; off = 0; acc = {k: 0 for k in ("ins", "del", "tr", "rot", "rei")}
    for step in range(160):
        nmol = eng.number_of_molecules(comp)
        kind = rng.choice(["ins", "del", "tr", "rot", "rei"])
        if kind in ("del", "tr", "rot", "rei") and nmol == 0:
            continue
        mol = int(rng.integers(0, max(nmol, 1)))
        if kind == "ins":
            fb = eng.cbmc_first_bead(CBMC_INSERTION, comp, 0, off, rng.random()); off += 10
            if not fb["success"]:
                continue
            ch = eng.cbmc_chain(CBMC_INSERTION, comp, 0, off, rng.random()); off += 10
            if not ch["success"]:
                continue
            ew = eng.ewald_delta(comp, INSERTION, location=ch["selected"])
            dE = fb["energy"].sum() + ch["energy"].sum() + ew.sum() + eng.tail_difference(comp, INSERTION)
            W = fb["rosenbluth"] * ch["rosenbluth"] * np.exp(-beta * ew.sum())
            if rng.random() < min(1.0, 0.02 * W):
                eng.accept_insertion(comp); run += dE; acc["ins"] += 1
        elif kind == "del":
            fb = eng.cbmc_first_bead(CBMC_DELETION, comp, mol, off, 0.5); off += 10
            ch = eng.cbmc_chain(CBMC_DELETION, comp, mol, off, 0.5); off += 10
            ew = eng.ewald_delta(comp, DELETION, location=mol * ms)
            dE = -(fb["energy"].sum() + ch["energy"].sum()) + ew.sum() + eng.tail_difference(comp, DELETION)
            if rng.random() < 0.3:
                eng.accept_deletion(comp, mol); run += dE; acc["del"] += 1
        elif kind in ("tr", "rot"):
            mt = TRANSLATION if kind == "tr" else ROTATION
            eng.single_body_propose(mt, comp, mol, (0.5, 0.5, 0.5), off, want_pos=False); off += 3
            d, ov = eng.single_body_delta(comp)
            if ov:
                continue
            ew = eng.ewald_delta(comp, mt)
            dE = sum(d[k] for k in ("HHVDW", "HGVDW", "GGVDW", "HHReal", "HGReal", "GGReal")) + ew.sum()
            if rng.random() < np.exp(min(0.0, -beta * dE)):
                eng.accept_translation(comp); run += dE; acc[kind] += 1
        else:
            fb = eng.cbmc_first_bead(REINSERTION_INSERTION, comp, mol, off, rng.random()); off += 10
            if not fb["success"]:
                continue
            ch = eng.cbmc_chain(REINSERTION_INSERTION, comp, mol, off, rng.random()); off += 10
            if not ch["success"]:
                continue
            eng.reinsertion_store(comp)
            rb = eng.cbmc_first_bead(REINSERTION_RETRACE, comp, mol, off, 0.5, stored_r=fb["stored_r"]); off += 1
            rc = eng.cbmc_chain(REINSERTION_RETRACE, comp, mol, off, 0.5); off += 10
            ew = eng.ewald_delta(comp, REINSERTION, location=mol * ms)
            dE = (fb["energy"].sum() + ch["energy"].sum()) - (rb["energy"].sum() + rc["energy"].sum()) + ew.sum()
            Wn = fb["rosenbluth"] * ch["rosenbluth"] * np.exp(-beta * ew.sum()); Wo = rb["rosenbluth"] * rc["rosenbluth"]
            if Wo > 0 and rng.random() < min(1.0, Wn / Wo):
                eng.accept_reinsertion(comp, mol); run += dE; acc["rei"] += 1
    E1 = _total(eng)
    assert sum(acc.values()) >= 20, acc
    scale = max(1.0, abs(E0), abs(E1))
    assert abs((E1 - E0) - run) < 1e-7 * scale, (E1 - E0, run, acc)
    # the stored structure factors followed the accepted moves: recomputing them changes nothing
    sa, _, _ = eng.download_structure_factors()
    eng.total_ewald(store=True)
    sb, _, _ = eng.download_structure_factors()
    assert np.max(np.abs(sa - sb)) < 1e-8
    eng.close()


def _same(a, b, tol=1e-11, floor=1e-6):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return bool(np.all(np.abs(a - b) <= tol * np.maximum(np.abs(b), floor) + tol * float(np.abs(b).max() if b.size else 0.0)))


def test_one_kernel_moves_match_the_stage_calls(gpu_engine_factory):
    """gb_move_* (one kernel, one host round trip) against the stage calls of the same move -- which the tests above pin to
    the oracle.  Same pool offsets and uniforms; energies agree to summation order (1e-11), selections exactly."""
    box, ff, s, z, eng = _setup(gpu_engine_factory)
    comp = 1; ms = 3
    rng = np.random.default_rng(31)
    pool = rng.random((1024, 3)); eng.upload_random_pool(pool)
    # ---- insertion
    for rep in range(8):
        off = 50 * rep; u = rng.random(2)
        fb = eng.cbmc_first_bead(CBMC_INSERTION, comp, 0, off, u[0])
        ch = eng.cbmc_chain(CBMC_INSERTION, comp, 0, off + 10, u[1]) if fb["success"] and fb["rosenbluth"] > 1e-150 else None
        ok = ch is not None and ch["success"] and fb["rosenbluth"] * ch["rosenbluth"] > 1e-150
        ew = eng.ewald_delta(comp, INSERTION, location=ch["selected"]) if ok else None
        grown = eng.cbmc_grown_positions(comp) if ok else None
        m = eng.move_insertion(comp, off, u)
        assert m["success"] == ok
        assert m["first_bead"]["success"] == fb["success"] and m["first_bead"]["n_survivors"] == fb["n_survivors"]
        if fb["success"] and fb["n_survivors"] > 0:
            assert m["first_bead"]["selected"] == fb["selected"]
            assert _same(m["first_bead"]["rosenbluth"], fb["rosenbluth"]) and _same(m["first_bead"]["energy"], fb["energy"])
        if ok:
            assert m["chain"]["selected"] == ch["selected"] and _same(m["chain"]["rosenbluth"], ch["rosenbluth"])
            assert _same(m["chain"]["energy"], ch["energy"])
            assert _same(m["ewald"], ew, tol=1e-10)
            assert np.allclose(eng.cbmc_grown_positions(comp), grown, rtol=0, atol=1e-12)
            assert m["uniforms_used"] == 2 and m["pool_used"] == 20
    # ---- deletion
    for mol in (0, 7, 19):
        off = 600 + 25 * (mol % 5)
        fb = eng.cbmc_first_bead(CBMC_DELETION, comp, mol, off, 0.5)
        ch = eng.cbmc_chain(CBMC_DELETION, comp, mol, off + 10, 0.5)
        ew = eng.ewald_delta(comp, DELETION, location=mol * ms)
        m = eng.move_deletion(comp, mol, off)
        assert m["success"] and m["first_bead"]["selected"] == 0 and m["uniforms_used"] == 0
        assert _same(m["first_bead"]["rosenbluth"], fb["rosenbluth"]) and _same(m["chain"]["rosenbluth"], ch["rosenbluth"])
        assert _same(m["first_bead"]["energy"], fb["energy"]) and _same(m["chain"]["energy"], ch["energy"])
        assert _same(m["ewald"], ew, tol=1e-10)
    # ---- reinsertion
    for mol, off in ((5, 200), (12, 300), (2, 420)):
        u = rng.random(2)
        fb = eng.cbmc_first_bead(REINSERTION_INSERTION, comp, mol, off, u[0])
        alive = fb["success"] and fb["rosenbluth"] > 1e-150
        ch = eng.cbmc_chain(REINSERTION_INSERTION, comp, mol, off + 10, u[1]) if alive else None
        alive = alive and ch["success"] and fb["rosenbluth"] * ch["rosenbluth"] > 1e-150
        m_ref = None
        if alive:
            eng.reinsertion_store(comp)
            rb = eng.cbmc_first_bead(REINSERTION_RETRACE, comp, mol, off + 20, 0.5, stored_r=fb["stored_r"])
            rc = eng.cbmc_chain(REINSERTION_RETRACE, comp, mol, off + 21, 0.5)
            ew = eng.ewald_delta(comp, REINSERTION, location=mol * ms)
            m_ref = (rb, rc, ew)
        m = eng.move_reinsertion(comp, mol, off, u)
        assert m["success"] == alive
        if alive:
            rb, rc, ew = m_ref
            assert m["first_bead"]["selected"] == fb["selected"] and m["chain"]["selected"] == ch["selected"]
            assert _same(m["first_bead"]["stored_r"], fb["stored_r"])
            assert _same(m["first_bead"]["rosenbluth"], fb["rosenbluth"]) and _same(m["chain"]["rosenbluth"], ch["rosenbluth"])
            assert _same(m["old_first_bead"]["rosenbluth"], rb["rosenbluth"]) and _same(m["old_chain"]["rosenbluth"], rc["rosenbluth"])
            assert _same(m["old_chain"]["energy"], rc["energy"])
            assert _same(m["ewald"], ew, tol=1e-10)
            assert m["pool_used"] == 31            # 10 positions + 10 orientations + 1 retrace position + 10 retrace orientations
    # ---- translation / rotation
    for k, (mt, mol) in enumerate([(TRANSLATION, 3), (ROTATION, 11), (TRANSLATION, 19), (ROTATION, 0)]):
        maxc = np.array([0.8, 0.6, 0.7]) if mt == TRANSLATION else np.array([0.5, 0.4, 0.3])
        eng.single_body_propose(mt, comp, mol, maxc, 900 + k)
        d, ov = eng.single_body_delta(comp)
        ew = eng.ewald_delta(comp, mt)
        m = eng.move_single_body(mt, comp, mol, maxc, 900 + k)
        assert m["overlap"] == bool(ov)
        got = np.array([m["delta"][k2] for k2 in ("HHVDW", "HHReal", "HGVDW", "HGReal", "GGVDW", "GGReal")])
        ref = np.array([d[k2] for k2 in ("HHVDW", "HHReal", "HGVDW", "HGReal", "GGVDW", "GGReal")])
        assert np.max(np.abs(got - ref)) < 1e-11 * 1e4
        if not ov:
            assert _same(m["ewald"], ew, tol=1e-10)
    eng.close()


def _totals(eng):
    v = eng.total_vdw_real(); w = eng.total_ewald(store=False)
    return (v["HHVDW"] + v["HGVDW"] + v["GGVDW"] + v["HHReal"] + v["HGReal"] + v["GGReal"]
            + w["HHEwaldE"] + w["HGEwaldE"] + w["GGEwaldE"] + eng.tail_total())


def test_identity_swap_stages_and_commit(gpu_engine_factory, oracle):
    """IdentitySwapMove (mc_swap_moves.h:199-431) on the Xe/Kr mixture (config D, no charges): IDENTITY_SWAP_NEW first
    bead at the old molecule's position with the old molecule excluded, IDENTITY_SWAP_OLD retrace, tail difference,
    then the commit: total energy recomputed from scratch must move by exactly the reported delta."""
    from graspa_b200.types import IDENTITY_SWAP_NEW, IDENTITY_SWAP_OLD, species_counts, pseudo_atom_counts
    box, ff, s, z, eng = _setup(gpu_engine_factory, "D", grow=None)
    beta = float(z["beta"])
    pool = np.random.default_rng(5).random((64, 3)); eng.upload_random_pool(pool)
    E0 = _totals(eng)
    running = 0.0
    nmol = {1: int(s.natoms[1]), 2: int(s.natoms[2])}
    cur = s
    for step, (oldc, newc, mol) in enumerate([(1, 2, 3), (2, 1, 0), (2, 2, 5), (1, 2, nmol[1] - 2)]):
        o = int(cur.offsets[oldc])
        old_pos = eng.download_atoms(oldc)["pos"][mol] if hasattr(eng, "download_atoms") else cur.pos[o + mol]
        fb = eng.cbmc_first_bead(IDENTITY_SWAP_NEW, newc, nmol[newc], step, 0.5, excl_comp=oldc, excl_mol=mol)
        assert fb["success"] and fb["uniform_used"] == 0 and fb["selected"] == 0
        assert np.allclose(fb["selected_pos"], old_pos, atol=0, rtol=0)
        eng.reinsertion_store(newc)
        rb = eng.cbmc_first_bead(IDENTITY_SWAP_OLD, oldc, mol, step, 0.5)
        assert rb["success"] and rb["selected"] == 0
        tail = eng.tail_identity_swap(newc, oldc)
        # the one-kernel move reports the same numbers as the three stage calls
        m = eng.move_identity_swap(oldc, mol, newc, step, 0.5)
        assert m["success"] and m["pool_used"] == 2 and m["uniforms_used"] == 0
        assert abs(m["first_bead"]["rosenbluth"] - fb["rosenbluth"]) <= 1e-12 * fb["rosenbluth"]          # summation order differs
        assert abs(m["old_first_bead"]["rosenbluth"] - rb["rosenbluth"]) <= 1e-12 * rb["rosenbluth"]
        assert _close(m["first_bead"]["energy"], fb["energy"]) and _close(m["old_first_bead"]["energy"], rb["energy"])
        assert m["tail"] == tail and np.array_equal(m["first_bead"]["selected_pos"], fb["selected_pos"])
        if step == 0:
            # oracle: one trial atom of the NEW species at the old position against the system minus the old molecule
            tr = TrialAtoms(np.array([old_pos]), np.array([cur.charge[int(cur.offsets[newc])]]), np.array([cur.type[int(cur.offsets[newc])]]))
            e, f, _ = oracle.trial_energies(box, ff, cur, 1, 1, tr, newc, nmol[newc], excl_comp=oldc, excl_mol=mol)
            assert _close(fb["energy"], e[0]) and abs(fb["rosenbluth"] - np.exp(-beta * e[0].sum())) <= 1e-9 * fb["rosenbluth"]
            tr_o = TrialAtoms(np.array([old_pos]), np.array([cur.charge[o]]), np.array([cur.type[o]]))
            eo, fo, _ = oracle.trial_energies(box, ff, cur, 1, 1, tr_o, oldc, mol)
            assert _close(rb["energy"], eo[0]) and abs(rb["rosenbluth"] - np.exp(-beta * eo[0].sum())) <= 1e-9 * rb["rosenbluth"]
            npseudo = pseudo_atom_counts(cur, ff.ntypes)
            t_ref = oracle.tail_identity_swap(ff, npseudo, box.volume, species_counts(cur, newc, ff.ntypes), species_counts(cur, oldc, ff.ntypes))
            assert abs(tail - t_ref) <= 1e-12 * max(1.0, abs(t_ref))
        delta = float(np.sum(fb["energy"]) - np.sum(rb["energy"]) + tail)
        eng.accept_identity_swap(oldc, mol, newc)
        if newc != oldc:
            nmol[newc] += 1; nmol[oldc] -= 1
        assert eng.number_of_molecules(newc) == nmol[newc] and eng.number_of_molecules(oldc) == nmol[oldc]
        running += delta
        E1 = _totals(eng)
        assert abs((E1 - E0) - running) <= 1e-9 * max(1.0, abs(E1)), (step, E1 - E0, running)
    eng.close()


def test_identity_swap_same_species_with_charges(gpu_engine_factory, oracle):
    """CO2 -> CO2 identity swap in CO2-MFI (config B): chain growth with the old molecule excluded, Ewald delta of
    GPU_EwaldDifference_IdentitySwap (old molecule out, stored new molecule in), commit, energy drift."""
    from graspa_b200.types import IDENTITY_SWAP_NEW, IDENTITY_SWAP_OLD
    box, ff, s, z, eng = _setup(gpu_engine_factory, "B")
    comp = 1; ms = 3; mol = 4
    pool = np.random.default_rng(6).random((64, 3)); eng.upload_random_pool(pool)
    eng.total_ewald(store=True)
    E0 = _totals(eng)
    o = int(s.offsets[comp]); nm = int(s.natoms[comp]) // ms
    fb = eng.cbmc_first_bead(IDENTITY_SWAP_NEW, comp, nm, 0, 0.5, excl_comp=comp, excl_mol=mol)
    assert fb["success"] and np.array_equal(fb["selected_pos"], s.pos[o + mol * ms])
    ch = eng.cbmc_chain(IDENTITY_SWAP_NEW, comp, nm, 1, 0.41, excl_comp=comp, excl_mol=mol)
    assert ch["success"] and ch["uniform_used"] == 1
    new_pos = eng.cbmc_grown_positions(comp)
    eng.reinsertion_store(comp)
    rb = eng.cbmc_first_bead(IDENTITY_SWAP_OLD, comp, mol, 11, 0.5)
    rc = eng.cbmc_chain(IDENTITY_SWAP_OLD, comp, mol, 12, 0.5)
    ew = eng.ewald_delta_identity_swap(comp, comp, mol * ms)
    q = s.charge[o:o + ms]
    pos = np.concatenate([s.pos[o + mol * ms:o + mol * ms + ms], new_pos])
    ref, _, _ = oracle.ewald_delta(box, pos, np.concatenate([q, q]), np.ones(2 * ms), ms, ms, z["sf_ads"], z["sf_fw"])
    assert _close(ew, ref, scale=max(1.0, float(np.abs(ref).max())))       # the two exclusion constants cancel (same species)
    delta = float(np.sum(fb["energy"]) + np.sum(ch["energy"]) - np.sum(rb["energy"]) - np.sum(rc["energy"]) + ew[0] + ew[1])
    # the one-kernel move (same pool layout: growth at 0 and 1.., retrace behind it) against the stage calls
    m = eng.move_identity_swap(comp, mol, comp, 0, 0.41)
    assert m["success"] and m["pool_used"] == 22 and m["uniforms_used"] == 1
    assert m["chain"]["selected"] == ch["selected"]
    for a, b in ((m["first_bead"], fb), (m["chain"], ch), (m["old_first_bead"], rb), (m["old_chain"], rc)):
        assert abs(a["rosenbluth"] - b["rosenbluth"]) <= 1e-12 * abs(b["rosenbluth"]) and _close(a["energy"], b["energy"])
    assert _close(np.array(m["ewald"]), ew, scale=max(1.0, float(np.abs(ew).max())))
    assert np.allclose(eng.snapshot_molecules(comp, 0, 1)["pos"], s.pos[o:o + ms])        # nothing committed yet
    eng.accept_identity_swap(comp, mol, comp)
    assert np.allclose(eng.snapshot_molecules(comp, mol, 1)["pos"], new_pos, atol=1e-12)  # the fused call left the same molecule in tempMolStorage
    E1 = _totals(eng)
    assert abs((E1 - E0) - delta) <= 1e-9 * max(1.0, abs(E1)), (E1 - E0, delta)
    eng.close()


def test_block_pockets_in_the_move_kernels(gpu_engine_factory, oracle):
    """gb_set_block_pockets: the device-side BlockedPocket test (read_data.cpp:3466-3640) as the kept drivers apply it --
    first-bead trials (a blocked starting bead flags every trial, mc_widom.h:445-497), the grown molecule after the chain
    stage (mc_swap_utilities.h:35-78), translation proposals (mc_single_particle.h:83-119) -- stage calls and fused
    calls against the oracle's flags."""
    box, ff, s, z, eng = _setup(gpu_engine_factory, "B")
    comp = 1; ms = 3; beta = float(z["beta"])
    rng = np.random.default_rng(31)
    pool = rng.random((4096, 3)); eng.upload_random_pool(pool)
    # pockets that cover roughly a third of the box
    cen = rng.random((40, 3)) * np.array([box.cell[0], box.cell[4], box.cell[8]]); rad = 3.0 + 3.0 * rng.random(40)
    eng.set_block_pockets(comp, cen, rad)
    blocked = lambda p: oracle.blocked_pocket(box, cen, rad, p)
    nmol = int(s.natoms[comp]) // ms
    seen = dict(all=0, some=0, none=0, grown=0)
    for k in range(60):
        off = 20 * k; u = 0.37
        tr = oracle.trial_positions(box, s, CBMC_INSERTION, comp, 0, 10, pool[off:off + 10])
        e, f, _ = oracle.trial_energies(box, ff, s, 10, 1, tr, comp, nmol)
        bl = np.array([blocked(p) for p in tr.pos])
        f2 = f.copy()
        if bl[0]: f2[:] = 1
        else: f2[1:] |= bl[1:].astype(f2.dtype)
        seen["all" if bl[0] else ("some" if bl.any() else "none")] += 1
        ref = oracle.cbmc_finish(CBMC_INSERTION, False, e, f2, beta, 10, u)
        for fused in (False, True):
            if fused:
                m = eng.move_insertion(comp, off, (u, 0.61)); fb = m["first_bead"]
            else:
                fb = eng.cbmc_first_bead(CBMC_INSERTION, comp, 0, off, u)
            assert fb["n_survivors"] == ref["nsurv"], (k, fused, fb, ref)
            assert fb["success"] == ref["success"]
            if not ref["success"]:
                continue
            assert fb["selected"] == ref["selected"] and abs(fb["rosenbluth"] - ref["rosenbluth"]) <= 1e-9 * ref["rosenbluth"]
            # chain stage: the grown molecule is checked atom by atom
            if fused:
                ch = m["chain"]; grown_ok = m["success"]
            else:
                ch = eng.cbmc_chain(CBMC_INSERTION, comp, 0, off + 10, 0.61); grown_ok = ch["success"]
            t2 = oracle.trial_orientations(s, CBMC_INSERTION, comp, 1, ms - 1, 10, pool[off + 10:off + 20], tr.pos[ref["selected"]])
            e2, fl2, _ = oracle.trial_energies(box, ff, s, 10, ms - 1, t2, comp, nmol)
            ref2 = oracle.cbmc_finish(CBMC_INSERTION, True, e2, fl2, beta, 10, 0.61)
            if ref2["success"] and ref["rosenbluth"] * ref2["rosenbluth"] > 1e-150:
                sel = ref2["selected"]
                mol = np.concatenate([[tr.pos[ref["selected"]]], t2.pos[sel * (ms - 1):(sel + 1) * (ms - 1)]])
                grown_blocked = any(blocked(p) for p in mol)
                seen["grown"] += int(grown_blocked)
                assert bool(grown_ok) == (not grown_blocked), (k, fused, grown_ok, grown_blocked)
            else:
                assert not grown_ok
    assert seen["all"] > 3 and seen["some"] > 3 and seen["grown"] > 0, seen
    # translation: a proposal with an atom inside a pocket is reported as an overlap, by both paths
    o = int(s.offsets[comp]); nb = 0
    for k in range(40):
        mol = k % nmol; off = 2000 + 3 * k
        new = eng.single_body_propose(TRANSLATION, comp, mol, (2.0, 2.0, 2.0), off)
        d, ov = eng.single_body_delta(comp)
        old = TrialAtoms(s.pos[o + mol * ms:o + mol * ms + ms], s.charge[o:o + ms], s.type[o:o + ms])
        nw = TrialAtoms(new, s.charge[o:o + ms], s.type[o:o + ms])
        _, ov_ref = oracle.single_body_delta(box, ff, s, comp, mol, old, nw)
        bl = any(blocked(p) for p in new)
        nb += int(bl)
        assert bool(ov) == bool(ov_ref or bl)
        m = eng.move_single_body(TRANSLATION, comp, mol, (2.0, 2.0, 2.0), off)
        assert bool(m["overlap"]) == bool(ov_ref or bl)
    assert nb > 3
    # removing the list restores the plain behaviour
    eng.set_block_pockets(comp, np.zeros((0, 3)), np.zeros(0))
    fb = eng.cbmc_first_bead(CBMC_INSERTION, comp, 0, 0, 0.37)
    tr = oracle.trial_positions(box, s, CBMC_INSERTION, comp, 0, 10, pool[0:10])
    e, f, _ = oracle.trial_energies(box, ff, s, 10, 1, tr, comp, nmol)
    assert fb["n_survivors"] == int((f == 0).sum())
    eng.close()


def test_isotherm_points_as_independent_boxes():
    """graspa_b200.boxes: one host-driver process per isotherm point, dealt over the visible GPUs, nothing exchanged.
    Loading must grow with pressure and every box must conserve energy (running sum vs recomputed totals)."""
    import os
    import torch
    from graspa_b200.boxes import run_boxes, DRIVER, ROOT
    deck = os.path.join(ROOT, "oracle", "_ref", "examples", "CO2-MFI")
    if not (os.path.exists(DRIVER) and os.path.isdir(deck)):
        pytest.skip("host driver / example deck not built")
    pts = [{"pressure": 1.0e3}, {"pressure": 1.0e5}]
    res, wall = run_boxes(deck, pts, gpus=max(1, min(2, torch.cuda.device_count())), init=300, prod=300)
    assert all(r["returncode"] == 0 for r in res), res
    lo, hi = res[0]["loading"][0]["production_average"], res[1]["loading"][0]["production_average"]
    assert hi > lo >= 0.0
    assert all(abs(r["energy_drift"]) < 1e-6 for r in res)
    assert abs(res[0]["pressure_pa"] - 1.0e3) < 1e-9 and abs(res[1]["pressure_pa"] - 1.0e5) < 1e-9


def test_cbcfc_lambda_change_vs_oracle(gpu_engine_factory, oracle):
    """CB/CFC lambda change of one molecule (CBCF_LambdaChange mc_cbcfc.h:20-140): VDW + real delta of
    Calculate_Single_Body_Energy_VDWReal_LambdaChange (VDW_Coulomb.cu:843-1036, soft-core LJ of maths.cuh:452-494), the
    Fourier delta of GPU_EwaldDifference_LambdaChange with its exclusion term (Ewald_Energy_Functions.h:637-788), and the
    commit: after accepting lambda = 0.35 the total energy recomputed from scratch moves by the reported delta; going on
    from there to lambda = 0.8 starts from the stored fractional scaling factors."""
    box, ff, s, z, eng = _setup(gpu_engine_factory, "B")
    comp = 1; ms = 3; mol = 6
    o = int(s.offsets[comp]); a0 = o + mol * ms
    eng.total_ewald(store=True)
    E0 = _totals(eng)
    excl = float(z["excl"][0]) + float(z["excl"][1])
    old_scale = (1.0, 1.0); running = 0.0
    cur = s
    for new_scale in ((0.35, 0.35 ** 5), (0.8, 0.8 ** 5), (1.0, 1.0)):
        d, ov = eng.lambda_change_delta(comp, mol, new_scale)
        pos = cur.pos[a0:a0 + ms]; q = cur.charge[a0:a0 + ms]; ty = cur.type[a0:a0 + ms]
        old = TrialAtoms(pos, q, ty, np.full(ms, old_scale[0]), np.full(ms, old_scale[1]))
        new = TrialAtoms(pos, q, ty, np.full(ms, new_scale[0]), np.full(ms, new_scale[1]))
        ref, ov_ref = oracle.single_body_delta(box, ff, cur, comp, mol, old, new)
        got = np.array([d["HHVDW"], d["HHReal"], d["HGVDW"], d["HGReal"], d["GGVDW"], d["GGReal"]])
        assert bool(ov) == bool(ov_ref)
        assert _close(got, ref, scale=max(1.0, float(np.abs(ref).max())))
        ew = eng.ewald_delta_lambda_change(comp, old_scale, new_scale)
        sa, sf, _ = eng.download_structure_factors()
        ref_ew, _, _ = oracle.ewald_delta(box, np.concatenate([pos, pos]), np.concatenate([q, q]),
                                          np.concatenate([np.full(ms, old_scale[1]), np.full(ms, new_scale[1])]), ms, ms, sa, sf)
        ref_ew[0] -= excl * (new_scale[1] ** 2 - old_scale[1] ** 2)
        assert _close(ew, ref_ew, scale=max(1.0, float(np.abs(ref_ew).max())))
        eng.accept_lambda_change(comp, mol, new_scale)
        running += float(got.sum() + ew[0] + ew[1])
        E1 = _totals(eng)
        assert abs((E1 - E0) - running) <= 1e-9 * max(1.0, abs(E1)), (new_scale, E1 - E0, running)
        # the oracle's copy of the system follows the commit
        sc = cur.scale.copy(); scc = cur.scale_coul.copy(); sc[a0:a0 + ms] = new_scale[0]; scc[a0:a0 + ms] = new_scale[1]
        from graspa_b200.types import System
        cur = System(cur.nhost, cur.natoms, cur.molsize, cur.pos, cur.charge, cur.type, cur.molid, sc, scc, alloc=cur.alloc)
        old_scale = new_scale
    assert abs(running) <= 1e-7 * max(1.0, abs(E0))          # back at lambda = 1: the three deltas cancel
    eng.close()


def _totals_no_tail(eng):
    v = eng.total_vdw_real(); w = eng.total_ewald(store=False)
    return (v["HHVDW"] + v["HGVDW"] + v["GGVDW"] + v["HHReal"] + v["HGReal"] + v["GGReal"] + w["HHEwaldE"] + w["HGEwaldE"] + w["GGEwaldE"])


def _sb_sum(d):
    return sum(d[k] for k in ("HHVDW", "HHReal", "HGVDW", "HGReal", "GGVDW", "GGReal"))


def test_cbcf_insertion_and_deletion_compositions(gpu_engine_factory, oracle):
    """The two-step CB/CFC moves of CBCFMove (mc_cbcfc.h:296-447) composed from the stage calls, with the intermediate Fourier state
    in tempEik (UseTempVector, Ewald_Energy_Functions.h:362-377, :510-517, :713-716).
    Insertion: the fractional molecule goes to lambda = 1 (update_CBCF_scale before the acceptance test), a new fractional molecule is
    grown at the new lambda and its Fourier delta continues from tempEik; accepted -> the from-scratch total energy moves by the sum of
    the two steps and the stored structure factors are those of the new state; rejected -> Revert_CBCF_Insertion leaves every slot and
    structure factor bit for bit as it was.
    Deletion: the fractional molecule is retraced and its Fourier delta written to tempEik, Update_deletion_data_fractional parks it
    behind the live range, another molecule goes from 1 to the new lambda on top of tempEik; accepted / rejected (Revert_CBCF_Deletion)
    as above.  The reference's own CBCFMove tests a local SuccessConstruction that nothing sets (mc_cbcfc.h:307-318, :372-393), so its
    program never accepts these branches: the energies are pinned here by the oracle and by the recomputed totals instead of a run."""
    box, ff, s, z, eng = _setup(gpu_engine_factory, "B", grow=600)
    comp = 1; ms = 3; o = int(s.offsets[comp])
    excl = float(z["excl"][0]) + float(z["excl"][1])
    rng = np.random.default_rng(77)
    pool = rng.random((4096, 3)); eng.upload_random_pool(pool)
    eng.total_ewald(store=True)
    lam = lambda x: (x, x ** 5)
    nmol0 = eng.number_of_molecules(comp)
    frac = 2; s_old = lam(0.4)
    eng.lambda_change_delta(comp, frac, s_old); eng.ewald_delta_lambda_change(comp, (1.0, 1.0), s_old); eng.accept_lambda_change(comp, frac, s_old)

    def insertion(accept, frac, s_old, s_new, off):
        E0 = _totals_no_tail(eng); a0 = eng.download_atoms(comp); sf0 = eng.download_structure_factors()[0].copy()
        d1, ov = eng.lambda_change_delta(comp, frac, (1.0, 1.0)); assert not ov
        ew1 = eng.ewald_delta_lambda_change(comp, s_old, (1.0, 1.0))
        eng.cbcf_set_scale(comp, frac, (1.0, 1.0))
        while True:
            fb = eng.cbmc_first_bead(CBMC_INSERTION, comp, 0, off, rng.random(), scale=s_new); off += 10
            if not fb["success"]:
                continue
            ch = eng.cbmc_chain(CBMC_INSERTION, comp, 0, off, rng.random()); off += 10
            if ch["success"]:
                break
        _, sfw, stmp = eng.download_structure_factors()
        ew2 = eng.ewald_delta(comp, CBCF_INSERTION, location=ch["selected"], scale=s_new)
        grown = eng.cbmc_grown_positions(comp)
        ref, _, _ = oracle.ewald_delta(box, grown, s.charge[o:o + ms], np.full(ms, s_new[1]), 0, ms, stmp, sfw)
        ref[0] -= excl * s_new[1] ** 2
        assert _close(ew2, ref, scale=max(1.0, float(np.abs(ref).max())))
        dE = _sb_sum(d1) + ew1.sum() + fb["energy"].sum() + ch["energy"].sum() + ew2.sum()
        if accept:
            eng.accept_insertion(comp)
            E1 = _totals_no_tail(eng)
            assert abs((E1 - E0) - dE) <= 1e-9 * max(1.0, abs(E1)), (E1 - E0, dE)
            n = eng.number_of_molecules(comp); a1 = eng.download_atoms(comp)
            assert n == a0["n_live"] // ms + 1
            assert np.all(a1["scale"][(n - 1) * ms:n * ms] == s_new[0]) and np.all(a1["scale_coul"][(n - 1) * ms:n * ms] == s_new[1])
            assert np.all(a1["scale"][frac * ms:(frac + 1) * ms] == 1.0) and np.all(a1["scale_coul"][frac * ms:(frac + 1) * ms] == 1.0)
            assert np.allclose(a1["pos"][(n - 1) * ms:n * ms], grown, rtol=0, atol=1e-12)
            sa = eng.download_structure_factors()[0].copy(); eng.total_ewald(store=True); sb = eng.download_structure_factors()[0]
            assert np.max(np.abs(sa - sb)) < 1e-8
        else:
            eng.cbcf_set_scale(comp, frac, s_old)
            a1 = eng.download_atoms(comp)
            for k in ("pos", "scale", "charge", "scale_coul", "type"):
                assert np.array_equal(a0[k][:a0["n_live"]], a1[k][:a1["n_live"]]), k
            assert a1["n_live"] == a0["n_live"] and np.array_equal(sf0, eng.download_structure_factors()[0])
            assert abs(_totals_no_tail(eng) - E0) <= 1e-12 * max(1.0, abs(E0))
        return off

    off = insertion(False, frac, s_old, lam(0.7), 0)
    off = insertion(True, frac, s_old, lam(0.7), off)
    assert eng.number_of_molecules(comp) == nmol0 + 1

    def deletion(accept, mdel, s_del, knew, s_new, off):
        E0 = _totals_no_tail(eng); a0 = eng.download_atoms(comp); sf0 = eng.download_structure_factors()[0].copy()
        n0 = eng.number_of_molecules(comp)
        fb = eng.cbmc_first_bead(CBMC_DELETION, comp, mdel, off, 0.5, scale=s_del); off += 10
        ch = eng.cbmc_chain(CBMC_DELETION, comp, mdel, off, 0.5); off += 10
        ewd = eng.ewald_delta(comp, CBCF_DELETION, location=mdel * ms, scale=s_del)
        eng.cbcf_deletion_stage(comp, mdel)
        assert eng.number_of_molecules(comp) == n0 - 1
        a_mid = eng.download_atoms(comp)
        if mdel != n0 - 1:        # the last molecule took the slot, the deleted one is parked behind the live range
            assert np.array_equal(a_mid["pos"][mdel * ms:(mdel + 1) * ms], a0["pos"][(n0 - 1) * ms:n0 * ms])
        d2, ov = eng.lambda_change_delta(comp, knew, s_new); assert not ov
        _, sfw, stmp = eng.download_structure_factors()
        ew2 = eng.ewald_delta_lambda_change(comp, (1.0, 1.0), s_new, use_temp_vector=True)
        pk = a_mid["pos"][knew * ms:(knew + 1) * ms]; q = s.charge[o:o + ms]
        ref, _, _ = oracle.ewald_delta(box, np.concatenate([pk, pk]), np.concatenate([q, q]),
                                       np.concatenate([np.ones(ms), np.full(ms, s_new[1])]), ms, ms, stmp, sfw)
        ref[0] -= excl * (s_new[1] ** 2 - 1.0)
        assert _close(ew2, ref, scale=max(1.0, float(np.abs(ref).max())))
        dE = -(fb["energy"].sum() + ch["energy"].sum()) + ewd.sum() + _sb_sum(d2) + ew2.sum()
        if accept:
            eng.accept_lambda_change(comp, knew, s_new)
            E1 = _totals_no_tail(eng)
            assert abs((E1 - E0) - dE) <= 1e-9 * max(1.0, abs(E1)), (E1 - E0, dE)
            sa = eng.download_structure_factors()[0].copy(); eng.total_ewald(store=True); sb = eng.download_structure_factors()[0]
            assert np.max(np.abs(sa - sb)) < 1e-8
        else:
            eng.cbcf_deletion_stage(comp, mdel, revert=True)
            a1 = eng.download_atoms(comp)
            for k in ("pos", "scale", "charge", "scale_coul", "type"):
                assert np.array_equal(a0[k][:a0["n_live"]], a1[k][:a1["n_live"]]), k
            assert a1["n_live"] == a0["n_live"] and np.array_equal(sf0, eng.download_structure_factors()[0])
            assert abs(_totals_no_tail(eng) - E0) <= 1e-12 * max(1.0, abs(E0))
        return off

    nfrac = eng.number_of_molecules(comp) - 1             # the molecule grown above is the fractional one now (lambda 0.7)
    off = deletion(False, nfrac, lam(0.7), 4, lam(0.3), off)
    # make a molecule in the middle the fractional one, so that the exchange with the last slot is exercised
    eng.lambda_change_delta(comp, nfrac, (1.0, 1.0)); eng.ewald_delta_lambda_change(comp, lam(0.7), (1.0, 1.0)); eng.accept_lambda_change(comp, nfrac, (1.0, 1.0))
    eng.lambda_change_delta(comp, 3, lam(0.6)); eng.ewald_delta_lambda_change(comp, (1.0, 1.0), lam(0.6)); eng.accept_lambda_change(comp, 3, lam(0.6))
    off = deletion(False, 3, lam(0.6), 5, lam(0.2), off)
    off = deletion(True, 3, lam(0.6), 5, lam(0.2), off)
    assert eng.number_of_molecules(comp) == nmol0
    eng.close()


def test_fractional_molecule_created_by_an_insertion_is_seen_by_later_moves(gpu_engine_factory):
    """CreateMolecule_InOneBox (axpy.cu:322-349) creates the fractional molecule of a CB/CFC component by an ordinary CBMC insertion
    with TempVal.Scale = SET_SCALE(lambda).  From then on every pair loop has to honour the scaling factors of system atoms (the engine
    skips them while every atom has scale 1): translations and reinsertions of the OTHER molecules against the recomputed totals.
    Found by the CB/CFC run of the bound reference program: one reinsertion next to the fresh fractional molecule was off by 0.049 in
    guest-guest VDW."""
    box, ff, s, z, eng = _setup(gpu_engine_factory, "B", grow=600)
    comp = 1; ms = 3
    rng = np.random.default_rng(5)
    pool = rng.random((8192, 3)); eng.upload_random_pool(pool)
    eng.total_ewald(store=True)
    E0 = _totals_no_tail(eng); run = 0.0; off = 0
    sc = (0.5, 0.5 ** 5)
    # grow the fractional molecule next to an existing one, so that guest-guest pairs with it matter
    target = s.pos[int(s.offsets[comp]) + 3 * 4]
    while True:
        fb = eng.cbmc_first_bead(CBMC_INSERTION, comp, 0, off, rng.random(), scale=sc); off += 10
        if not fb["success"]:
            continue
        ch = eng.cbmc_chain(CBMC_INSERTION, comp, 0, off, rng.random()); off += 10
        if ch["success"]:
            break
    ew = eng.ewald_delta(comp, INSERTION, location=ch["selected"], scale=sc)
    eng.accept_insertion(comp); run += fb["energy"].sum() + ch["energy"].sum() + ew.sum()
    nmol = eng.number_of_molecules(comp); frac = nmol - 1
    a = eng.download_atoms(comp)
    assert np.all(a["scale"][frac * ms:(frac + 1) * ms] == sc[0]) and np.all(a["scale_coul"][frac * ms:(frac + 1) * ms] == sc[1])
    assert abs((_totals_no_tail(eng) - E0) - run) <= 1e-9 * max(1.0, abs(E0))
    # bring molecules close to the fractional one and move them around it
    done = 0
    for mol in range(nmol - 1):
        for kind in (TRANSLATION, ROTATION):
            eng.single_body_propose(kind, comp, mol, (1.0, 1.0, 1.0), off, want_pos=False); off += 3
            d, ov = eng.single_body_delta(comp)
            if ov:
                continue
            ew = eng.ewald_delta(comp, kind)
            eng.accept_translation(comp); run += _sb_sum(d) + ew.sum(); done += 1
        fb = eng.cbmc_first_bead(REINSERTION_INSERTION, comp, mol, off, rng.random()); off += 10
        if not fb["success"]:
            continue
        ch = eng.cbmc_chain(REINSERTION_INSERTION, comp, mol, off, rng.random()); off += 10
        if not ch["success"]:
            continue
        eng.reinsertion_store(comp)
        rb = eng.cbmc_first_bead(REINSERTION_RETRACE, comp, mol, off, 0.5, stored_r=fb["stored_r"]); off += 1
        rc = eng.cbmc_chain(REINSERTION_RETRACE, comp, mol, off, 0.5); off += 10
        ew = eng.ewald_delta(comp, REINSERTION, location=mol * ms)
        eng.accept_reinsertion(comp, mol); done += 1
        run += (fb["energy"].sum() + ch["energy"].sum()) - (rb["energy"].sum() + rc["energy"].sum()) + ew.sum()
    E1 = _totals_no_tail(eng)
    assert done >= nmol
    assert abs((E1 - E0) - run) <= 1e-9 * max(1.0, abs(E0), abs(E1)), (E1 - E0, run)
    eng.close()


@pytest.mark.parametrize("b", [1, 4])
def test_host_driver_reads_and_writes_raspa2_restarts(b, tmp_path):
    """The NIST SPC/E decks end to end through the C++ host driver: `RestartFile yes` (RestartFileParser,
    read_data.cpp:3000-3221), zero cycles, the printed initial energies against the reference's own output.txt values
    (tests/golden/nist_spce.npz), then the snapshot written by --write-restart read back: same molecules."""
    import os
    import subprocess
    from graspa_b200.boxes import DRIVER, ROOT
    from tests.conftest import load_nist
    from tests.support.raspa_inputs import read_restart_positions
    deck = os.path.join(ROOT, "oracle", "_ref", "examples", "Reference_NIST_SPCE", f"Box-{b}")
    if not (os.path.exists(DRIVER) and os.path.isdir(deck)):
        pytest.skip("host driver / NIST decks not built (oracle/build_ref.sh examples)")
    out = tmp_path / "restartfile"
    r = subprocess.run([DRIVER, deck, "--write-restart", str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    box, ff, s, ref = load_nist(b)
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("INITIAL")][0]
    val = lambda key: float(line.split(key + ":")[1].split(",")[0])
    assert abs(val("VDW [Guest-Guest]") - ref["vdw_gg"]) < 2e-5
    assert abs(val("Real [Guest-Guest]") - ref["real_gg"]) < 2e-5
    assert abs(val("Ewald [Guest-Guest]") - ref["ewald_gg"]) < 2e-5
    assert abs(val("Tail") - ref["tail"]) < 2e-5
    assert abs(val("Total") - ref["total"]) < 5e-5
    n = int(s.natoms[1])
    pos, chg = read_restart_positions(str(out), 0, 3, box, with_charge=True)
    assert pos.shape == (n, 3)
    assert np.max(np.abs(pos - s.pos[:n])) < 1e-9 and np.max(np.abs(chg - s.charge[:n])) < 1e-12


def _scaled_state(box, s, scale, dk=0):
    """what ScalePositions leaves (mc_box.h:18-94): cell * scale, every molecule of components >= 1 moved with its first atom,
    kmax grown by dk so that the number of wave vectors really changes; -> (new Box, new System)"""
    from graspa_b200.types import Box, System
    kmax = tuple(int(k) + dk for k in box.kmax)
    nb = Box(box.cell * scale, alpha=box.alpha, kmax=kmax, recip_cutoff=(1.05 * max(kmax)) ** 2, prefactor=box.prefactor)
    from oracle import oracle as orc
    pos = orc.scale_positions(box, s, scale)          # ScalePositions restated in oracle/oracle.py
    return nb, System(s.nhost, s.natoms.copy(), s.molsize.copy(), pos, s.charge.copy(), s.type.copy(), s.molid.copy(), alloc=s.alloc.copy())


@pytest.mark.parametrize("scale,dk", [(1.013, 1), (0.991, 0)])
def test_npt_volume_move_vs_oracle(gpu_engine_factory, oracle, scale, dk):
    """gb_volume_move_trial / _finish (VolumeMove, mc_box.h:196-320): energies of the scaled system against the oracle's totals
    on the restated ScalePositions result, rejection restores positions, box and structure factors exactly, acceptance leaves a
    state on which the next move's deltas are those of the new box."""
    box, ff, s, z, eng = _setup(gpu_engine_factory)
    comp = 1; ms = 3; o = int(s.offsets[comp]); nmol = int(s.natoms[comp]) // ms
    E0 = eng.total_ewald(store=True); V0 = eng.total_vdw_real()
    sa0, sf0, _ = eng.download_structure_factors()
    p0 = eng.snapshot_molecules(comp, 0, nmol)["pos"]
    nb, ns = _scaled_state(box, s, scale, dk)

    def check_trial():
        got, ov = eng.volume_move_trial(nb, scale)
        ref = oracle.total_vdw_real(nb, ff, ns)       # HHv, HHr, HGv, HGr, GGv, GGr
        g = np.array([got["HHVDW"], got["HHReal"], got["HGVDW"], got["HGReal"], got["GGVDW"], got["GGReal"]])
        assert np.max(np.abs(g - ref)) < ETOL * max(1.0, float(np.abs(ref).max()))
        Ew, sa, sf = oracle.ewald_total(nb, ns)       # GG (incl. HH), HH, HG as the CPU routine reports them
        dev = np.array([Ew[0] - Ew[1], Ew[1], Ew[2]])   # the device routine keeps HH out of GG
        ge = np.array([got["GGEwaldE"], got["HHEwaldE"], got["HGEwaldE"]])
        assert np.max(np.abs(ge - dev)) < ETOL * max(1.0, float(np.abs(dev).max()))
        assert ov == 0
        assert np.allclose(eng.snapshot_molecules(comp, 0, nmol)["pos"], ns.pos[o:o + nmol * ms], rtol=0, atol=1e-11)
        return sa, sf

    check_trial()
    with pytest.raises(Exception):
        eng.upload_box(box)                            # nothing else may change the state while the move is pending
    eng.volume_move_finish(False)
    # rejected: bitwise the state before the trial
    assert np.array_equal(eng.snapshot_molecules(comp, 0, nmol)["pos"], p0)
    sa1, sf1, _ = eng.download_structure_factors()
    assert np.array_equal(sa1, sa0) and np.array_equal(sf1, sf0)
    assert eng.total_vdw_real() == V0 and eng.total_ewald(store=False) == E0
    # accepted: stored structure factors are those of the new state and the next move's deltas belong to the new box
    sa_ref, sf_ref = check_trial()
    eng.volume_move_finish(True)
    sa2, sf2, _ = eng.download_structure_factors()
    mag = max(1.0, float(np.abs(sf_ref).max()))
    assert np.max(np.abs(sa2 - sa_ref)) < 1e-9 * mag and np.max(np.abs(sf2 - sf_ref)) < 1e-9 * mag
    assert abs(eng.tail_total() - oracle.tail_total(ff, pseudo_atom_counts(ns, ff.ntypes), nb.volume)) < 1e-9
    rng = np.random.default_rng(5)
    pool = rng.random((8, 3)); eng.upload_random_pool(pool)
    mol = 7; maxc = np.array([0.6, 0.7, 0.5])
    newp = eng.single_body_propose(TRANSLATION, comp, mol, maxc, 0)
    oldp = ns.pos[o + mol * ms:o + mol * ms + ms]
    assert np.allclose(newp, oldp + maxc * 2.0 * (pool[0] - 0.5), rtol=0, atol=1e-11)
    d, ov = eng.single_body_delta(comp)
    q = s.charge[o:o + ms]; ty = s.type[o:o + ms]
    ref, rov = oracle.single_body_delta(nb, ff, ns, comp, mol, TrialAtoms(oldp, q, ty), TrialAtoms(newp, q, ty))
    got = np.array([d["HHVDW"], d["HHReal"], d["HGVDW"], d["HGReal"], d["GGVDW"], d["GGReal"]])
    assert ov == rov and np.max(np.abs(got - ref)) < ETOL * 1e4
    ew = eng.ewald_delta(comp, TRANSLATION)
    ewr, _, _ = oracle.ewald_delta(nb, np.concatenate([oldp, newp]), np.concatenate([q, q]), np.ones(2 * ms), ms, ms, sa_ref, sf_ref)
    assert _close(ew, ewr, scale=max(1.0, float(np.abs(ewr).max())))
    eng.close()


def test_npt_volume_move_reports_overlap(gpu_engine_factory):
    """a strong compression pushes molecules into each other / the framework: the overlap flag of Total_VDW_Coulomb_Energy"""
    box, ff, s, z, eng = _setup(gpu_engine_factory)
    nb, ns = _scaled_state(box, s, 0.70)
    _, ov = eng.volume_move_trial(nb, 0.70)
    assert ov == 1
    eng.volume_move_finish(False)
    eng.close()


def _run_driver(deck_name, *flags):
    import os
    import subprocess
    from graspa_b200.boxes import DRIVER, ROOT
    deck = os.path.join(ROOT, "oracle", "_ref", "examples", deck_name)
    if not (os.path.exists(DRIVER) and os.path.isdir(deck)):
        pytest.skip("host driver / example deck not built (oracle/build_ref.sh examples, make -C graspa_b200/host)")
    r = subprocess.run([DRIVER, deck, *flags], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout.splitlines()


def _term(line, key):
    return float(line.split(key + ":")[1].split(",")[0])


def test_host_driver_npt_example_matches_the_reference_run():
    """Examples/NPTMC (100 CO2 created in an empty box, NPT volume moves), 1500 + 500 cycles, seed 0: the numbers below are the
    reference CUDA program's own for the same deck and seed (profiles/r1_trace_parity.txt: 202 000 moves, no decision differs)"""
    out = _run_driver("NPTMC", "--init", "1500", "--equil", "0", "--prod", "500")
    final = [ln for ln in out if ln.startswith("FINAL")][0]
    assert abs(_term(final, "VDW [Guest-Guest]") + 6116.32502) < 2e-5 and abs(_term(final, "Real [Guest-Guest]") + 2450.30981) < 2e-5
    vol = [ln for ln in out if ln.startswith("Volume Move:")][0]
    assert "4704/6448 accepted" in vol
    cyc = [ln for ln in out if ln.startswith("CYCLE:")]
    assert cyc[1] == "CYCLE: 500, AccRatio: 0.83496, compare_to_target_ratio: 1.50000, MaxVolumeChange: 0.05625"
    drift = float([ln for ln in out if ln.startswith("ENERGY DRIFT")][0].split(":")[-1])
    assert abs(drift) < 1e-6


def test_host_driver_gibbs_example_matches_the_reference_run():
    """Examples/NVT-Gibbs (two boxes run together, 852 + 151 CO2, particle transfers and volume exchange), 30 + 20 cycles, seed 0:
    counters and final energies of both boxes as the reference CUDA program prints them; molecules and volume are conserved"""
    out = _run_driver("NVT-Gibbs", "--init", "30", "--equil", "0", "--prod", "20")
    finals = [ln for ln in out if ln.startswith("FINAL")]
    assert len(finals) == 2
    assert abs(_term(finals[0], "VDW [Guest-Guest]") + 669378.85067) < 2e-5 and abs(_term(finals[0], "Real [Guest-Guest]") + 218849.56576) < 2e-5
    assert abs(_term(finals[1], "VDW [Guest-Guest]") + 3399.58198) < 2e-5 and abs(_term(finals[1], "Real [Guest-Guest]") + 1465.16641) < 2e-5
    g = [ln for ln in out if ln.startswith("Gibbs Volume Move:")][0]
    assert "Gibbs Volume Move: 139/1286 accepted" in g and "Gibbs Particle Transfer: 3228/13332 accepted" in g
    bv = [ln for ln in out if ln.startswith("Box volumes:")][0]
    v0, v1 = (float(x) for x in bv.split(":")[1].split(";")[0].split())
    assert abs(v0 + v1 - (39.0 ** 3 + 71.0 ** 3)) < 1e-6
    n0, n1 = (int(x) for x in bv.split("molecules:")[1].split()[0].split("/"))
    assert n0 + n1 == 852 + 151
    drifts = [float(ln.split(":")[-1]) for ln in out if ln.startswith("ENERGY DRIFT")]
    assert len(drifts) == 2 and all(abs(d) < 1e-5 for d in drifts)
