"""Parity of the CUDA engine (through the C ABI) against the oracle and the committed golden vectors.
Tolerances are BASELINE.json's: energies 1e-10 relative, Widom <W> 1e-9 relative."""
import numpy as np
import pytest

from graspa_b200.types import TrialAtoms, species_counts, pseudo_atom_counts, INSERTION, DELETION
from tests.conftest import load_config, rel_err

pytestmark = pytest.mark.gpu
CONFIGS = ["A", "E", "B", "C", "D"]
ETOL = 1e-10


def _scale(e):
    return np.abs(e).sum(axis=1, keepdims=True) + 1e-3


@pytest.fixture(params=["warp", "cells"])
def widom_path(request, monkeypatch):
    """both pair stages of gb_widom_batch: the warp-per-insertion kernel (k_widom_pair, small batches) and the cell-sorted one
    (widom_cells.cuh, batches with >= 64 first-bead trials per cell); GB_WIDOM_PATH forces either for any batch size"""
    monkeypatch.setenv("GB_WIDOM_PATH", request.param)
    return request.param


@pytest.mark.parametrize("name", CONFIGS)
def test_trial_energies_vs_golden(gpu_engine_factory, name):
    box, ff, s, z = load_config(name)
    eng = gpu_engine_factory(box, ff, s)
    comp = int(z["comp"]); ms = int(s.molsize[comp]); new_molid = int(s.natoms[comp]) // ms
    t1 = TrialAtoms(z["tb1_pos"], z["tb1_charge"], z["tb1_type"])
    e, f = eng.trial_energies(len(z["tb1_flag"]), 1, t1, comp, new_molid)
    assert (f == z["tb1_flag"]).all()
    assert np.max(np.abs(e - z["tb1_energy"]) / _scale(z["tb1_energy"])) < ETOL
    if "tb2_pos" in z:
        cs = int(z["tb2_cs"])
        t2 = TrialAtoms(z["tb2_pos"], z["tb2_charge"], z["tb2_type"])
        e, f = eng.trial_energies(len(z["tb2_flag"]), cs, t2, comp, new_molid)
        assert (f == z["tb2_flag"]).all()
        assert np.max(np.abs(e - z["tb2_energy"]) / _scale(z["tb2_energy"])) < ETOL
    eng.close()


def test_trial_energies_edge_cases(gpu_engine_factory, oracle):
    """exclusion of a molecule (deletion/retrace), a trial on top of an atom (r^2 < 0.01 flag), ragged sizes"""
    box, ff, s, z = load_config("B")
    eng = gpu_engine_factory(box, ff, s)
    comp = 1; o = int(s.offsets[comp]); ms = 3
    # retrace of existing molecule 4: its own atoms must be excluded
    mol = 4
    tr = TrialAtoms(s.pos[o + ms * mol:o + ms * mol + 1], s.charge[o + ms * mol:o + ms * mol + 1], s.type[o + ms * mol:o + ms * mol + 1])
    e_gpu, f_gpu = eng.trial_energies(1, 1, tr, comp, mol)
    e_cpu, f_cpu, _ = oracle.trial_energies(box, ff, s, 1, 1, tr, comp, mol)
    assert (f_gpu == f_cpu).all() and np.max(np.abs(e_gpu - e_cpu) / _scale(e_cpu)) < ETOL
    # same position, nothing excluded: sits on top of itself -> overlap flag through r^2 < 0.01
    e_gpu, f_gpu = eng.trial_energies(1, 1, tr, comp, 999)
    e_cpu, f_cpu, _ = oracle.trial_energies(box, ff, s, 1, 1, tr, comp, 999)
    assert f_gpu[0] == 1 and f_cpu[0] == 1
    # ExcludeList[0] (identity swap): exclude molecule 7 of component 1
    rng = np.random.default_rng(5)
    pos = s.pos[o + ms * 7:o + ms * 7 + 1] + 0.3
    tr = TrialAtoms(pos, [0.0], [int(s.type[o])])
    e_gpu, f_gpu = eng.trial_energies(1, 1, tr, comp, 999, excl_comp=1, excl_mol=7)
    e_cpu, f_cpu, _ = oracle.trial_energies(box, ff, s, 1, 1, tr, comp, 999, excl_comp=1, excl_mol=7)
    assert (f_gpu == f_cpu).all() and np.max(np.abs(e_gpu - e_cpu) / _scale(e_cpu)) < ETOL
    # ragged: 37 trials of a 5-atom group
    n = 37 * 5
    tr = TrialAtoms(rng.random((n, 3)) * 30.0, rng.normal(size=n) * 0.3, rng.integers(0, ff.ntypes, size=n))
    e_gpu, f_gpu = eng.trial_energies(37, 5, tr, comp, 999)
    e_cpu, f_cpu, _ = oracle.trial_energies(box, ff, s, 37, 5, tr, comp, 999)
    assert (f_gpu == f_cpu).all() and np.max(np.abs(e_gpu - e_cpu) / _scale(e_cpu)) < ETOL
    eng.close()


@pytest.mark.parametrize("name", ["A", "E", "B", "C"])
def test_ewald_total_and_structure_factors(gpu_engine_factory, name):
    box, ff, s, z = load_config(name)
    eng = gpu_engine_factory(box, ff, s)
    E = eng.total_ewald(store=True)
    ref = z["ewald_E"]          # GG, HH, HG as the reference's Ewald_Total returns them
    got = np.array([E["GGEwaldE"], E["HHEwaldE"], E["HGEwaldE"]])
    assert rel_err(got, ref, floor=max(1.0, float(np.abs(ref).max()) * 1e-3)) < ETOL
    sa, sf, _ = eng.download_structure_factors()
    mag = max(1.0, float(np.abs(z["sf_fw"]).max()))
    assert np.max(np.abs(sa - z["sf_ads"])) < 1e-10 * mag * 10 and np.max(np.abs(sf - z["sf_fw"])) < 1e-10 * mag * 10
    eng.close()


@pytest.mark.parametrize("name", ["A", "B", "C"])
def test_ewald_delta_vs_oracle(gpu_engine_factory, oracle, name):
    box, ff, s, z = load_config(name)
    eng = gpu_engine_factory(box, ff, s)
    eng.upload_structure_factors(z["sf_ads"], z["sf_fw"])
    rng = np.random.default_rng(11)
    comp = int(z["comp"]); o = int(s.offsets[comp]); ms = int(s.molsize[comp])
    q = s.charge[o:o + ms]
    # insertion (0 old, ms new), deletion (ms old, 0 new), translation (ms old + ms new)
    newp = rng.random((ms, 3)) * 20.0; oldp = rng.random((ms, 3)) * 20.0
    for nold, nnew, pos, qq in [(0, ms, newp, q), (ms, 0, oldp, q), (ms, ms, np.concatenate([oldp, newp]), np.concatenate([q, q]))]:
        got = eng.ewald_delta_explicit(False, nold, nnew, pos, qq, np.ones(len(qq)))
        ref, temp, _ = oracle.ewald_delta(box, pos, qq, np.ones(len(qq)), nold, nnew, z["sf_ads"], z["sf_fw"])
        mag = max(1.0, float(np.abs(ref).max()))
        assert np.max(np.abs(got - ref)) / mag < ETOL
        _, _, tgpu = eng.download_structure_factors()
        assert np.max(np.abs(tgpu - temp)) < 1e-9
    # commit = pointer swap: the adsorbate structure factors become the temp ones
    eng.ewald_commit(comp)
    sa, _, _ = eng.download_structure_factors()
    assert np.max(np.abs(sa - temp)) < 1e-9
    eng.close()


@pytest.mark.parametrize("name", CONFIGS)
def test_tail(gpu_engine_factory, name):
    box, ff, s, z = load_config(name)
    eng = gpu_engine_factory(box, ff, s)
    assert (eng.pseudo_atom_counts() == z["npseudo"]).all()
    assert abs(eng.tail_total() - float(z["tail_total"])) <= 1e-12 * max(1.0, abs(float(z["tail_total"])))
    for k, c in enumerate(range(s.nhost, s.ncomp)):
        assert abs(eng.tail_difference(c, INSERTION) - z["tail_ins"][k]) <= 1e-12 * max(1.0, abs(z["tail_ins"][k]))
        assert abs(eng.tail_difference(c, DELETION) - z["tail_del"][k]) <= 1e-12 * max(1.0, abs(z["tail_del"][k]))
    if "tail_swap" in z:
        assert abs(eng.tail_identity_swap(s.nhost, s.nhost + 1) - float(z["tail_swap"])) <= 1e-12 * max(1.0, abs(float(z["tail_swap"])))
    eng.close()


@pytest.mark.parametrize("name", ["B", "C", "D"])
def test_total_vdw_real_vs_oracle(gpu_engine_factory, oracle, name):
    box, ff, s, z = load_config(name)
    eng = gpu_engine_factory(box, ff, s)
    got = eng.total_vdw_real()
    ref = oracle.total_vdw_real(box, ff, s)       # HHv, HHr, HGv, HGr, GGv, GGr
    g = np.array([got["HHVDW"], got["HHReal"], got["HGVDW"], got["HGReal"], got["GGVDW"], got["GGReal"]])
    assert rel_err(g, ref, floor=max(1.0, float(np.abs(ref).max()) * 1e-6)) < ETOL
    eng.close()


@pytest.mark.parametrize("name", CONFIGS)
def test_widom_batch_vs_golden(gpu_engine_factory, name, widom_path):
    box, ff, s, z = load_config(name)
    comp = int(z["comp"])
    eng = gpu_engine_factory(box, ff, s, float(z["beta"]), int(z["ntrials"]), int(z["norient"]))
    if "sf_ads" in z:
        eng.upload_structure_factors(z["sf_ads"], z["sf_fw"])
        eng.set_exclusion_constants(comp, float(z["excl"][0]), float(z["excl"][1]))
    rnd = z["widom_rnd"].reshape(-1, 3)
    out, stage, sums = eng.widom_batch(comp, rnd, z["widom_uni"], n_blocks=5)
    ref = z["widom_out"]
    # the engine distinguishes "chain failed" (2) from "chain failed with no surviving orientation" (3); the oracle reports 2 for both
    assert (np.where(stage == 3, 2, stage) == z["widom_stage"]).all()
    assert rel_err(out[:, 0], ref[:, 0], floor=1e-290) < 1e-9                   # W
    esc = np.abs(ref[:, 1:]).sum(axis=1, keepdims=True) + 1e-3
    assert np.max(np.abs(out[:, 1:] - ref[:, 1:]) / esc) < ETOL                 # energy terms
    # block sums = RecordRosen + widom_energy accumulation
    assert abs(sums[:, 0].sum() - ref[:, 0].sum()) <= 1e-9 * abs(ref[:, 0].sum())
    assert abs(sums[:, 1].sum() - (ref[:, 0] ** 2).sum()) <= 1e-9 * (ref[:, 0] ** 2).sum()
    assert sums[:, 2].sum() == len(ref)
    assert np.allclose(sums[:, 3:10].sum(axis=0), (ref[:, :1] * ref[:, 1:]).sum(axis=0), rtol=1e-8, atol=1e-6)
    eng.close()


def test_widom_batch_vs_the_reference_program(gpu_engine_factory, widom_path):
    """gb_widom_batch against numbers the REFERENCE PROGRAM wrote while it ran (tests/golden/ref_dump_widom_A.npz: the first 150
    Widom insertions of Examples/Henrys_coefficient, seed 0, from the instrumented reference CUDA build): same pool randoms and
    uniforms in, final Rosenbluth weight within 1e-9 relative and every energy term out."""
    import os
    from tests.conftest import GOLDEN
    box, ff, s, z = load_config("A")
    d = dict(np.load(os.path.join(GOLDEN, "ref_dump_widom_A.npz")))
    comp = int(z["comp"])
    eng = gpu_engine_factory(box, ff, s, float(z["beta"]), 10, 10)
    eng.upload_structure_factors(z["sf_ads"], z["sf_fw"]); eng.set_exclusion_constants(comp, float(z["excl"][0]), float(z["excl"][1]))
    rnd = np.concatenate([d["fb_rnd"], d["ch_rnd"]], axis=1)
    uni = np.stack([d["fbs"][:, 3], d["chs"][:, 3]], axis=1)
    out, stage, sums = eng.widom_batch(comp, rnd.reshape(-1, 3), uni)
    ref = d["ins"]
    assert (stage == 0).all() and d["has_ins"].all()
    assert np.max(np.abs(out[:, 0] - ref[:, 0]) / ref[:, 0]) < 1e-9
    esc = np.abs(ref[:, 1:]).sum(axis=1, keepdims=True) + 1e-3
    assert np.max(np.abs(out[:, 1:] - ref[:, 1:]) / esc) < 1e-9
    eng.close()


def test_widom_batch_vs_oracle_larger_with_failures(gpu_engine_factory, oracle, widom_path):
    """1024 insertions in config A with OverlapCriteria lowered so that first beads and chains fail:
    identical stage codes, W within 1e-9, and explicit pool indices (the RNG-exact replay layout)."""
    box, ff, s, z = load_config("A")
    ff.overlap = 2.0e4
    comp = 1; n = 1024
    rng = np.random.default_rng(99)
    rnd = rng.random((n, 20, 3)); uni = rng.random((n, 2))
    ws = oracle.WidomSetup(box, ff, s, comp, float(z["beta"]), 10, 10, z["sf_ads"], z["sf_fw"])
    ref, rstage, _ = oracle.widom_batch(ws, rnd, uni)
    eng = gpu_engine_factory(box, ff, s, float(z["beta"]), 10, 10)
    eng.upload_structure_factors(z["sf_ads"], z["sf_fw"]); eng.set_exclusion_constants(comp, *ws.excl)
    out, stage, sums = eng.widom_batch(comp, rnd.reshape(-1, 3), uni)
    # the engine distinguishes "chain failed" (2) from "chain failed with no surviving orientation" (3); the oracle reports 2 for both
    assert (np.where(stage == 3, 2, stage) == rstage).all()
    assert (rstage > 0).sum() > 0
    assert rel_err(out[:, 0], ref[:, 0], floor=1e-290) < 1e-9
    assert abs(sums[:, 0].sum() / n - ref[:, 0].mean()) <= 1e-9 * ref[:, 0].mean()
    assert sums[:, 10].sum() == (rstage > 0).sum()
    # shuffled pool with explicit indices gives the same answers
    perm = rng.permutation(2 * n)
    pool = np.zeros((2 * n * 10, 3))
    blocks = rnd.reshape(2 * n, 10, 3)
    for b in range(2 * n):
        pool[perm[b] * 10:(perm[b] + 1) * 10] = blocks[b]
    fb = perm[0::2] * 10; orr = perm[1::2] * 10
    out2, stage2, _ = eng.widom_batch(comp, pool, uni, fb_index=fb, or_index=orr)
    assert (stage2 == stage).all() and np.array_equal(out2, out)
    eng.close()


def test_widom_properties_full_size(gpu_engine_factory, widom_path):
    """size-independent properties at the benchmark shape (config E): determinism, batch-split additivity,
    block sums consistent with per-insertion outputs, W >= 0 and stage/W consistency"""
    box, ff, s, z = load_config("E")
    comp = 1; n = 20000
    rng = np.random.default_rng(3)
    rnd = rng.random((n * 20, 3)); uni = rng.random((n, 2))
    eng = gpu_engine_factory(box, ff, s, float(z["beta"]), 10, 10)
    eng.upload_structure_factors(z["sf_ads"], z["sf_fw"]); eng.set_exclusion_constants(comp, float(z["excl"][0]), float(z["excl"][1]))
    out, stage, sums = eng.widom_batch(comp, rnd, uni)
    out_b, stage_b, sums_b = eng.widom_batch(comp, rnd, uni)
    assert np.array_equal(out, out_b) and np.array_equal(sums, sums_b)            # bitwise deterministic
    h = n // 2
    o1, s1, _ = eng.widom_batch(comp, rnd[:h * 20], uni[:h]); o2, s2, _ = eng.widom_batch(comp, rnd[h * 20:], uni[h:])
    assert np.array_equal(np.concatenate([o1, o2]), out)                          # insertions are independent
    assert (out[:, 0] >= 0).all() and ((out[:, 0] == 0) == (stage > 0)).all()
    assert abs(sums[:, 0].sum() - out[:, 0].sum()) <= 1e-11 * out[:, 0].sum()
    assert sums[:, 2].sum() == n and (sums[:, 2] == n // 5).all()
    eng.close()


def test_widom_shards_bin_on_the_global_index(gpu_engine_factory, widom_path):
    """a job cut into contiguous index ranges (gb_widom_inputs.global_first/global_n) gives, after summing the
    per-shard block sums (what the NCCL all-reduce does), the single-call sums: counts exactly, sums to association."""
    from graspa_b200.shard import shard_range
    box, ff, s, z = load_config("A")
    comp = int(z["comp"]); n = 1003
    eng = gpu_engine_factory(box, ff, s, float(z["beta"]), 10, 10)
    eng.upload_structure_factors(z["sf_ads"], z["sf_fw"]); eng.set_exclusion_constants(comp, float(z["excl"][0]), float(z["excl"][1]))
    rng = np.random.default_rng(5)
    rnd = rng.random((n, 20, 3)); uni = rng.random((n, 2))
    out, stage, whole = eng.widom_batch(comp, rnd.reshape(-1, 3), uni, n_blocks=5)
    for world in (2, 3, 8):
        acc = np.zeros_like(whole); outs = []
        for r in range(world):
            first, cnt = shard_range(n, world, r)
            o, st, sm = eng.widom_batch(comp, rnd[first:first + cnt].reshape(-1, 3), uni[first:first + cnt], n_blocks=5, shard=(first, n))
            acc += sm; outs.append(o)
        assert np.array_equal(np.concatenate(outs), out)                          # per-insertion results do not depend on the cut
        assert np.array_equal(acc[:, 2], whole[:, 2]) and np.array_equal(acc[:, 10], whole[:, 10])
        assert np.max(np.abs(acc - whole) / np.maximum(np.abs(whole), 1e-300)) < 1e-12
    with pytest.raises(Exception):
        eng.widom_batch(comp, rnd[:10].reshape(-1, 3), uni[:10], shard=(n - 5, n))   # range sticks out of the job
    eng.close()


@pytest.mark.parametrize("b", [1, 2, 3, 4])
def test_nist_spce_known_answers(gpu_engine_factory, b):
    """The reference's known-answer example (Examples/Reference_NIST_SPCE): total energies of 100-400 SPC/E waters in
    triclinic boxes through the C ABI against the reference's printed output.txt values (5 decimals), LAMMPS-style
    Ewald set-up, an empty framework component, O-O tail correction."""
    from tests.conftest import load_nist
    box, ff, s, ref = load_nist(b)
    eng = gpu_engine_factory(box, ff, s)
    v = eng.total_vdw_real(); E = eng.total_ewald(store=True); tail = eng.tail_total()
    assert v["HHVDW"] == 0.0 and v["HGVDW"] == 0.0 and v["HHReal"] == 0.0 and v["HGReal"] == 0.0
    assert abs(v["GGVDW"] - ref["vdw_gg"]) <= 1e-10 * abs(ref["vdw_gg"]) + 1e-5
    assert abs(v["GGReal"] - ref["real_gg"]) <= 1e-10 * abs(ref["real_gg"]) + 1e-5
    assert abs(E["GGEwaldE"] - ref["ewald_gg"]) <= 1e-10 * abs(ref["ewald_gg"]) + 1e-5
    assert E["HHEwaldE"] == 0.0 and E["HGEwaldE"] == 0.0
    assert abs(tail - ref["tail"]) <= 1e-5
    total = v["GGVDW"] + v["GGReal"] + E["GGEwaldE"] + tail
    assert abs(total - ref["total"]) <= 2e-5
    eng.close()


def _as_1264(ff, seed=3):
    """the same Lennard-Jones table restated in the polynomial form UseLJ1264 uses (read_data.cpp:1196-1230:
    C12 = 4 eps sigma^12, C6 = 4 eps sigma^6), plus r^-4 and r^-10 terms of a size that matters at adsorption distances"""
    from graspa_b200.types import ForceField
    rng = np.random.default_rng(seed)
    n = ff.ntypes
    s6 = ff.sigma ** 6
    c4 = rng.normal(size=(n, n)) * 30.0; c4 = 0.5 * (c4 + c4.T)
    c10 = rng.normal(size=(n, n)) * 3.0e4; c10 = 0.5 * (c10 + c10.T)
    C12 = 4.0 * ff.epsilon * s6 * s6; C6 = 4.0 * ff.epsilon * s6
    ri2 = 1.0 / ff.cutoff_vdw_sq
    shift = C12 * ri2 ** 6 - C6 * ri2 ** 3 + c10.ravel() * ri2 ** 5 + c4.ravel() * ri2 ** 2      # Get_Shifted_Value_Coeff read_data.cpp:750-758
    return ForceField(C12, C6, shift, ff.cutoff_vdw, ff.cutoff_coul, overlap=ff.overlap, no_charges=ff.no_charges, use1264=True,
                      z=c4.ravel(), c10=c10.ravel())


@pytest.mark.parametrize("name", ["A", "B"])
def test_polynomial_1264_potential_vs_oracle(gpu_engine_factory, oracle, name, widom_path):
    """UseLJ1264: U = C12/r^12 - C6/r^6 + C10/r^10 + C4/r^4 - shift (VDW, maths.cuh:452-476) through trial energies,
    whole-system totals and a Widom batch"""
    box, ff0, s, z = load_config(name)
    ff = _as_1264(ff0)
    comp = int(z["comp"]); ms = int(s.molsize[comp])
    eng = gpu_engine_factory(box, ff, s, float(z["beta"]), 10, 10)
    rng = np.random.default_rng(17)
    L = np.array([box.cell[0], box.cell[4], box.cell[8]])
    n = 64 * 3
    tr = TrialAtoms(rng.random((n, 3)) * L, rng.normal(size=n) * 0.3, rng.integers(0, ff.ntypes, size=n))
    e_gpu, f_gpu = eng.trial_energies(64, 3, tr, comp, 999)
    e_cpu, f_cpu, _ = oracle.trial_energies(box, ff, s, 64, 3, tr, comp, 999)
    assert (f_gpu == f_cpu).all()
    e_gpu, f_gpu = eng.trial_energies(n, 1, tr, comp, 999)            # single atoms: enough of them survive OverlapCriteria
    e_cpu, f_cpu, _ = oracle.trial_energies(box, ff, s, n, 1, tr, comp, 999)
    assert (f_gpu == f_cpu).all() and (f_cpu == 0).sum() > 30
    ok = f_cpu == 0
    assert np.max(np.abs(e_gpu[ok] - e_cpu[ok]) / _scale(e_cpu[ok])) < ETOL
    # the r^-4 / r^-10 terms are really in play
    e_lj, f_lj, _ = oracle.trial_energies(box, ff0, s, n, 1, tr, comp, 999)
    both = ok & (f_lj == 0)
    assert np.max(np.abs(e_lj[both] - e_cpu[both])) > 1.0
    got = eng.total_vdw_real()
    ref = oracle.total_vdw_real(box, ff, s)
    g = np.array([got["HHVDW"], got["HHReal"], got["HGVDW"], got["HGReal"], got["GGVDW"], got["GGReal"]])
    assert rel_err(g, ref, floor=max(1.0, float(np.abs(ref).max()) * 1e-6)) < ETOL
    nw = 256
    rnd = rng.random((nw, 20, 3)); uni = rng.random((nw, 2))
    ws = oracle.WidomSetup(box, ff, s, comp, float(z["beta"]), 10, 10, z["sf_ads"], z["sf_fw"])
    wref, rstage, _ = oracle.widom_batch(ws, rnd, uni)
    eng.upload_structure_factors(z["sf_ads"], z["sf_fw"]); eng.set_exclusion_constants(comp, *ws.excl)
    out, stage, sums = eng.widom_batch(comp, rnd.reshape(-1, 3), uni)
    assert (np.where(stage == 3, 2, stage) == rstage).all()
    assert rel_err(out[:, 0], wref[:, 0], floor=1e-290) < 1e-9
    eng.close()


def test_widom_fourier_row_walk_equals_the_flat_loop(gpu_engine_factory, oracle):
    """k_widom_ewald walks the active wave vectors by (kx, ky) rows (a lane per row, |kz| in lockstep); the one-k-per-lane loop
    it replaced stays for molecules of more than four atoms and behind GB_EWALD_FLAT.  Same W and Fourier terms from both,
    and from the oracle, in a triclinic (A) and an orthorhombic box with adsorbates present (B)."""
    import os
    for name in ("A", "B"):
        box, ff, s, z = load_config(name)
        comp = int(z["comp"]); n = 512
        rng = np.random.default_rng(31)
        rnd = rng.random((n, 20, 3)); uni = rng.random((n, 2))
        ws = oracle.WidomSetup(box, ff, s, comp, float(z["beta"]), 10, 10, z["sf_ads"], z["sf_fw"])
        ref, rstage, _ = oracle.widom_batch(ws, rnd, uni)
        eng = gpu_engine_factory(box, ff, s, float(z["beta"]), 10, 10)
        eng.upload_structure_factors(z["sf_ads"], z["sf_fw"]); eng.set_exclusion_constants(comp, *ws.excl)
        out_rows, st_rows, _ = eng.widom_batch(comp, rnd.reshape(-1, 3), uni)
        os.environ["GB_EWALD_FLAT"] = "1"
        try:
            out_flat, st_flat, _ = eng.widom_batch(comp, rnd.reshape(-1, 3), uni)
        finally:
            del os.environ["GB_EWALD_FLAT"]
        assert (st_rows == st_flat).all()
        ok = st_rows == 0
        assert ok.sum() > 50
        # columns 5, 6: the same-species and cross Fourier terms of the insertion
        scale = np.abs(out_flat[ok][:, 5:7]).max()
        assert np.max(np.abs(out_rows[ok][:, 5:7] - out_flat[ok][:, 5:7])) < 1e-11 * max(1.0, scale)
        assert not np.array_equal(out_rows[ok][:, 5:7], out_flat[ok][:, 5:7]) or scale == 0.0     # two different summation orders really ran
        assert rel_err(out_rows[ok][:, 0], out_flat[ok][:, 0], floor=1e-290) < 1e-10
        assert rel_err(out_rows[:, 0], ref[:, 0], floor=1e-290) < 1e-9
        eng.close()


def test_widom_host_batches_are_pipelined_without_changing_results(gpu_engine_factory, widom_path):
    """Host-input batches of >= 65 536 insertions go up in chunks on a copy stream while the pair kernel already runs on what has
    arrived (gb_widom_batch); per-insertion results, stage codes and block sums are bitwise those of the single-copy path
    (GB_WIDOM_NO_OVERLAP=1) and of other chunk counts (ragged chunk boundaries included)."""
    import os
    box, ff, s, z = load_config("A")
    comp = 1; n = 70001                                  # not a multiple of anything the chunking rounds to
    rng = np.random.default_rng(8)
    rnd = rng.random((n * 20, 3)); uni = rng.random((n, 2))
    eng = gpu_engine_factory(box, ff, s, float(z["beta"]), 10, 10)
    eng.upload_structure_factors(z["sf_ads"], z["sf_fw"]); eng.set_exclusion_constants(comp, float(z["excl"][0]), float(z["excl"][1]))
    out, stage, sums = eng.widom_batch(comp, rnd, uni)
    for key, val in (("GB_WIDOM_NO_OVERLAP", "1"), ("GB_WIDOM_CHUNKS", "5"), ("GB_WIDOM_CHUNKS", "8")):
        os.environ[key] = val
        try:
            o2, s2, m2 = eng.widom_batch(comp, rnd, uni)
        finally:
            del os.environ[key]
        assert np.array_equal(o2, out) and np.array_equal(s2, stage) and np.array_equal(m2, sums), (key, val)
    assert (stage == 0).sum() > n // 2 and sums[:, 2].sum() == n
    eng.close()


def test_widom_paths_agree_and_large_batches_take_the_cell_sorted_stage(gpu_engine_factory, monkeypatch):
    """the two pair stages give the same insertions (energies to summation order, identical stage codes), with adsorbates present
    (config B: guest-guest terms) and without (E); an unforced batch of 100 000 takes the cell-sorted stage, one of 20 000 does not"""
    for name in ("E", "B", "C"):
        box, ff, s, z = load_config(name)
        comp = int(z["comp"]); n = 100000 if name == "E" else 3000
        rng = np.random.default_rng(77)
        rnd = rng.random((n * 20, 3)); uni = rng.random((n, 2))
        eng = gpu_engine_factory(box, ff, s, float(z["beta"]), 10, 10)
        eng.upload_structure_factors(z["sf_ads"], z["sf_fw"]); eng.set_exclusion_constants(comp, float(z["excl"][0]), float(z["excl"][1]))
        res = {}
        for path in ("warp", "cells"):
            monkeypatch.setenv("GB_WIDOM_PATH", path)
            res[path] = eng.widom_batch(comp, rnd, uni)
        monkeypatch.delenv("GB_WIDOM_PATH")
        (ow, sw, mw), (oc, sc, mc) = res["warp"], res["cells"]
        assert np.array_equal(sw, sc)
        ok = sw == 0
        assert ok.sum() > n // 4
        assert rel_err(oc[ok][:, 0], ow[ok][:, 0], floor=1e-290) < 1e-10
        esc = np.abs(ow[:, 1:]).sum(axis=1, keepdims=True) + 1e-3
        assert np.max(np.abs(oc[:, 1:] - ow[:, 1:]) / esc) < 1e-11
        assert not np.array_equal(oc[ok][:, 1:5], ow[ok][:, 1:5])                    # two different summation orders really ran
        if name == "E":
            o3, s3, m3 = eng.widom_batch(comp, rnd, uni)                              # unforced: 79 first-bead trials per cell (12 600 cells)
            assert np.array_equal(o3, oc) and np.array_equal(m3, mc)
            o4, s4, m4 = eng.widom_batch(comp, rnd[:20000 * 20], uni[:20000])         # unforced, 16 per cell: the warp-per-insertion kernel
            assert np.array_equal(o4, ow[:20000])
        eng.close()


def test_widom_batch_honours_block_pockets(gpu_engine_factory):
    """Config C with its block pockets: the batched path applies the first-bead rule (a blocked starting bead flags every trial,
    mc_widom.h:463-498) and the grown-molecule rule (mc_swap_utilities.h:46-80) like gb_move_insertion, which
    test_block_pockets_in_the_move_kernels pins to the oracle; same pool blocks and uniforms through both"""
    box, ff, s, z = load_config("C")
    comp = int(z["comp"]); n = 400; beta = float(z["beta"])
    # the deck's 8 pockets plus wide ones so that a good share of the insertions is touched by the rule
    rng = np.random.default_rng(91)
    L = np.array([box.cell[0], box.cell[4], box.cell[8]])
    centers = np.concatenate([z["pocket_centers"], rng.random((24, 3)) * L]); radii = np.concatenate([z["pocket_radii"], np.full(24, 3.5)])
    eng = gpu_engine_factory(box, ff, s, beta, 10, 10)
    eng.upload_structure_factors(z["sf_ads"], z["sf_fw"]); eng.set_exclusion_constants(comp, float(z["excl"][0]), float(z["excl"][1]))
    rnd = rng.random((n * 20, 3)); uni = rng.random((n, 2))
    eng.set_block_pockets(comp, centers, radii, invert=bool(z["pocket_invert"]))
    out, stage, sums = eng.widom_batch(comp, rnd, uni)
    eng.set_block_pockets(comp, np.zeros((0, 3)), np.zeros(0))
    out_free, stage_free, _ = eng.widom_batch(comp, rnd, uni)
    assert (stage != stage_free).sum() > 10                                          # the pockets really change outcomes
    eng.set_block_pockets(comp, centers, radii, invert=bool(z["pocket_invert"]))
    eng.upload_random_pool(rnd)
    tail = eng.tail_difference(comp, INSERTION)
    nfail = 0
    for i in range(n):
        m = eng.move_insertion(comp, 20 * i, uni[i])
        if not m["success"]:
            nfail += 1
            assert stage[i] != 0 and out[i, 0] == 0.0, (i, stage[i])
            continue
        W = m["first_bead"]["rosenbluth"] * m["chain"]["rosenbluth"] * np.exp(-beta * (m["ewald"][0] + m["ewald"][1])) * np.exp(-beta * tail)
        assert stage[i] == 0 and abs(out[i, 0] - W) <= 1e-9 * W, (i, out[i, 0], W)
    assert nfail > 10 and nfail < n
    eng.close()


def test_widom_replay_resumes_from_the_kept_first_bead_energies(gpu_engine_factory):
    """the RNG-exact replay of the host driver (mc_driver.cpp run_widom_batched_v2): every pool block is classified once
    (gb_widom_first_bead_success on the pool gb_upload_random_pool left on the device), the walk pairs first-bead and orientation
    blocks, and gb_widom_batch(resume_first_bead) starts from the kept first-bead energies.  Same insertions, to the bit, as a batch
    that is handed the pool and evaluates its first beads itself; a call that may change the system in between invalidates the kept
    energies (GB_ERR_STATE)."""
    from graspa_b200.engine import EngineError
    box, ff, s, z = load_config("A")
    comp = int(z["comp"])
    eng = gpu_engine_factory(box, ff, s, float(z["beta"]), 10, 10)
    eng.upload_structure_factors(z["sf_ads"], z["sf_fw"]); eng.set_exclusion_constants(comp, float(z["excl"][0]), float(z["excl"][1]))
    rng = np.random.default_rng(404)
    nblk = 16000                                                       # 160 000 first-bead trials: >= 16 per cell of config A, the kept-energies path
    pool = rng.random((nblk * 10, 3)); eng.upload_random_pool(pool)
    blocks = np.arange(nblk, dtype=np.int64) * 10
    code = eng.widom_first_bead_success(comp, None, blocks)
    ref_code = eng.widom_first_bead_success(comp, pool, blocks)          # handed the pool: nothing is kept
    assert np.array_equal(code, ref_code) and (code == 1).sum() > nblk // 4
    with pytest.raises(EngineError):
        eng.widom_batch(comp, None, np.full((1, 2), 0.5), fb_index=[0], or_index=[10], resume=True)
    eng.upload_random_pool(pool)
    code = eng.widom_first_bead_success(comp, None, blocks)
    # the driver's walk: a successful first bead consumes the next block as its orientation block
    fb, orr, k = [], [], 0
    while k + 1 < nblk:
        if code[k] == 1:
            fb.append(k * 10); orr.append((k + 1) * 10); k += 2
        else:
            fb.append(k * 10); orr.append(k * 10); k += 1
    uni = rng.random((len(fb), 2))
    o1, s1, m1 = eng.widom_batch(comp, None, uni, fb_index=fb, or_index=orr, n_blocks=1, resume=True)
    o2, s2, m2 = eng.widom_batch(comp, None, uni, fb_index=fb, or_index=orr, n_blocks=1, resume=True)       # nothing but Widom calls since: still valid
    o0, s0, m0 = eng.widom_batch(comp, pool, uni, fb_index=fb, or_index=orr, n_blocks=1)
    assert np.array_equal(s1, s0) and np.array_equal(s2, s0)
    assert (s0 == 0).sum() > len(fb) // 8
    # the unresumed batch of this size takes the warp-per-insertion kernel (another summation order; the replay's cells also follow
    # the pool: wider than 2 A for this small one): same selections, energies to rounding; the replay itself is bitwise repeatable
    esc = np.abs(o0[:, 1:]).sum(axis=1, keepdims=True) + 1e-3
    assert np.max(np.abs(o1[:, 1:] - o0[:, 1:]) / esc) < 1e-11 and rel_err(o1[s0 == 0][:, 0], o0[s0 == 0][:, 0], floor=1e-290) < 1e-10
    assert np.array_equal(o1, o2) and np.array_equal(m1, m2)
    # a GPU that holds only a share of the pool's blocks (the multi-GPU replay): the same bits for the insertions of that share
    half = nblk // 2
    eng.upload_random_pool(pool)
    code_h = eng.widom_first_bead_success(comp, None, blocks[half:])
    assert np.array_equal(code_h, code[half:])
    first = int(np.searchsorted(np.asarray(fb), half * 10))
    oh, sh, mh = eng.widom_batch(comp, None, uni[first:], fb_index=fb[first:], or_index=orr[first:], n_blocks=1, resume=True)
    assert np.array_equal(oh, o1[first:]) and np.array_equal(sh, s1[first:])
    with pytest.raises(EngineError):                                   # a first bead outside the share this engine classified
        eng.widom_batch(comp, None, uni[:4], fb_index=fb[:4], or_index=orr[:4], n_blocks=1, resume=True)
    o0, s0, m0 = eng.widom_batch(comp, pool, uni, fb_index=fb, or_index=orr, n_blocks=1)
    # the explicit pool replaced the engine's: the kept energies belong to another pool generation
    with pytest.raises(EngineError):
        eng.widom_batch(comp, None, uni, fb_index=fb, or_index=orr, n_blocks=1, resume=True)
    eng.close()


@pytest.mark.parametrize("name,ngroups", [("E", 280000), ("B", 120000), ("C", 150000)])
def test_large_trial_batches_take_the_cell_sorted_kernel(gpu_engine_factory, monkeypatch, name, ngroups):
    """gb_trial_energies with a batch of >= 64 trial atoms per 2 A cell (equal groups, nothing excluded) is evaluated by the cell-sorted
    energy kernel of the Widom stage; same energies (to summation order) and the same overlap flags as one CTA per group"""
    from graspa_b200.types import TrialAtoms
    box, ff, s, z = load_config(name)
    comp = int(z["comp"]); ms = int(s.molsize[comp]); o = int(s.offsets[comp])
    eng = gpu_engine_factory(box, ff, s)
    rng = np.random.default_rng(12)
    cell = box.cell.reshape(3, 3)
    tmpl = s.pos[o:o + ms] - s.pos[o]
    first = rng.random((ngroups, 3)) @ cell
    ang = rng.random(ngroups) * 2 * np.pi
    R = np.zeros((ngroups, 3, 3)); R[:, 0, 0] = np.cos(ang); R[:, 0, 1] = -np.sin(ang); R[:, 1, 0] = np.sin(ang); R[:, 1, 1] = np.cos(ang); R[:, 2, 2] = 1.0
    pos = (first[:, None, :] + np.einsum("nij,aj->nai", R, tmpl)).reshape(-1, 3)
    tr = TrialAtoms(pos, np.tile(s.charge[o:o + ms], ngroups), np.tile(s.type[o:o + ms], ngroups))
    n0 = eng.launch_count()
    e1, f1 = eng.trial_energies(ngroups, ms, tr, comp, 10 ** 9)
    assert eng.launch_count() - n0 >= 6                                  # pack, bin, scan, scatter, energy, group sums
    monkeypatch.setenv("GB_TRIAL_NO_CELLS", "1")
    n0 = eng.launch_count()
    e0, f0 = eng.trial_energies(ngroups, ms, tr, comp, 10 ** 9)
    assert eng.launch_count() - n0 == 1
    assert np.array_equal(f0, f1) and (f0 == 0).sum() > ngroups // 20
    ok = f0 == 0
    assert np.max(np.abs(e1[ok] - e0[ok]) / _scale(e0[ok])) < 1e-11
    assert not np.array_equal(e1[ok], e0[ok])
    eng.close()
