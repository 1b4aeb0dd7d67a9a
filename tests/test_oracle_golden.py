"""The plain-C oracle against the golden vectors generated from the reference's own routines
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from graspa_b200.types import TrialAtoms, species_counts, pseudo_atom_counts
from tests.conftest import load_config, rel_err

CONFIGS = ["A", "E", "B", "C", "D"]


@pytest.mark.parametrize("name", CONFIGS)
def test_trial_energies_match_reference(oracle, name):
    box, ff, s, z = load_config(name)
    comp = int(z["comp"]); ms = int(s.molsize[comp]); new_molid = int(s.natoms[comp]) // ms
    t1 = TrialAtoms(z["tb1_pos"], z["tb1_charge"], z["tb1_type"])
    e, f, c = oracle.trial_energies(box, ff, s, len(z["tb1_flag"]), 1, t1, comp, new_molid)
    assert (f == z["tb1_flag"]).all()
    assert (c[:3] == z["tb1_counts"]).all()
    # tolerance: 1e-10 relative (BASELINE.json), against the per-trial magnitude
    scale = np.abs(z["tb1_energy"]).sum(axis=1, keepdims=True) + 1e-3
    assert np.max(np.abs(e - z["tb1_energy"]) / scale) < 1e-10
    if "tb2_pos" in z:
        cs = int(z["tb2_cs"])
        t2 = TrialAtoms(z["tb2_pos"], z["tb2_charge"], z["tb2_type"])
        e, f, c = oracle.trial_energies(box, ff, s, len(z["tb2_flag"]), cs, t2, comp, new_molid)
        assert (f == z["tb2_flag"]).all() and (c[:3] == z["tb2_counts"]).all()
        scale = np.abs(z["tb2_energy"]).sum(axis=1, keepdims=True) + 1e-3
        assert np.max(np.abs(e - z["tb2_energy"]) / scale) < 1e-10


@pytest.mark.parametrize("name", ["A", "E", "B", "C"])
def test_ewald_total_and_structure_factors(oracle, name):
    box, ff, s, z = load_config(name)
    E, sa, sf = oracle.ewald_total(box, s)
    assert rel_err(E, z["ewald_E"], floor=1.0) < 1e-10
    assert np.max(np.abs(sa - z["sf_ads"])) < 1e-9 and np.max(np.abs(sf - z["sf_fw"])) < 1e-9
    comp = int(z["comp"]); o = int(s.offsets[comp]); ms = int(s.molsize[comp])
    ex = oracle.exclusion_rigid(box, s.pos[o:o + ms], s.charge[o:o + ms], s.scale_coul[o:o + ms])
    assert rel_err(ex, z["excl"]) < 1e-12


@pytest.mark.parametrize("name", CONFIGS)
def test_tail(oracle, name):
    box, ff, s, z = load_config(name)
    npseudo = pseudo_atom_counts(s, ff.ntypes)
    assert (npseudo == z["npseudo"]).all()
    assert abs(oracle.tail_total(ff, npseudo, box.volume) - float(z["tail_total"])) <= 1e-12 * max(1.0, abs(float(z["tail_total"])))
    for k, c in enumerate(range(s.nhost, s.ncomp)):
        cnt = species_counts(s, c, ff.ntypes)
        assert abs(oracle.tail_difference(ff, npseudo, box.volume, cnt, +1) - z["tail_ins"][k]) <= 1e-12 * max(1.0, abs(z["tail_ins"][k]))
        assert abs(oracle.tail_difference(ff, npseudo, box.volume, cnt, -1) - z["tail_del"][k]) <= 1e-12 * max(1.0, abs(z["tail_del"][k]))
    if "tail_swap" in z:
        a = oracle.tail_identity_swap(ff, npseudo, box.volume, species_counts(s, s.nhost, ff.ntypes), species_counts(s, s.nhost + 1, ff.ntypes))
        assert abs(a - float(z["tail_swap"])) <= 1e-12 * max(1.0, abs(float(z["tail_swap"])))
    if name == "D":
        assert float(z["tail_total"]) != 0.0      # the O-O override of force_field.def is active


@pytest.mark.parametrize("name", CONFIGS)
def test_widom_insertions_reproduce(oracle, name):
    box, ff, s, z = load_config(name)
    comp = int(z["comp"])
    ws = oracle.WidomSetup(box, ff, s, comp, float(z["beta"]), int(z["ntrials"]), int(z["norient"]), z.get("sf_ads"), z.get("sf_fw"))
    out, stage, counts = oracle.widom_batch(ws, z["widom_rnd"], z["widom_uni"], nthreads=2)
    assert (stage == z["widom_stage"]).all()
    assert rel_err(out[:, 0], z["widom_out"][:, 0], floor=1e-300) < 1e-12
    assert (counts == z["widom_counts"]).all()


def _moved(s, z, k):
    c = int(z["sb_comp"][k]); mol = int(z["sb_mol"][k]); ms = int(s.molsize[c]); o = int(s.offsets[c])
    sl = slice(o + mol * ms, o + (mol + 1) * ms)
    old = TrialAtoms(z["sb_old"][k][:3 * ms], s.charge[sl], s.type[sl]); new = TrialAtoms(z["sb_new"][k][:3 * ms], s.charge[sl], s.type[sl])
    return c, mol, ms, old, new


@pytest.mark.parametrize("name", ["B", "C", "D"])
def test_single_body_delta_matches_reference(oracle, name):
    """orc_single_body_delta (Calculate_Single_Body_Energy_VDWReal, VDW_Coulomb.cu:626-841, with its HH / HG / GG block layout,
    mc_single_particle.h:150-165) against new - old through the REFERENCE's pair routines (fixture sb_delta): adsorbate molecules
    in B, C, D and the movable Na+ of framework component 1 in C (cubic PBC branch, shifted potentials)."""
    box, ff, s, z = load_config(name)
    seen_host = False
    for k in range(len(z["sb_comp"])):
        c, mol, ms, old, new = _moved(s, z, k)
        seen_host |= c < s.nhost
        d, flag = oracle.single_body_delta(box, ff, s, c, mol, old, new)
        ref = z["sb_delta"][k]
        assert flag == int(z["sb_flag"][k])
        assert np.max(np.abs(d - ref)) <= 1e-10 * max(1.0, float(np.abs(ref).max())), (name, k, d, ref)
    assert seen_host == (name == "C")


@pytest.mark.parametrize("name", ["B", "C"])
def test_ewald_delta_matches_difference_of_reference_totals(oracle, name):
    """orc_ewald_delta (Fourier_Ewald_Diff, Ewald_Energy_Functions.h:280-397) against Ewald_Total(after) - Ewald_Total(before) of
    the REFERENCE (fixture sb_ewald), and its temp vector against the reference's structure factors of the moved state.  For a
    framework component (C, Na+) the same-type vector is FrameworkEik (:460-467)."""
    box, ff, s, z = load_config(name)
    for k in range(len(z["sb_comp"])):
        c, mol, ms, old, new = _moved(s, z, k)
        host = c < s.nhost
        same, cross = (z["sf_fw"], z["sf_ads"]) if host else (z["sf_ads"], z["sf_fw"])
        pos = np.concatenate([old.pos, new.pos]); q = np.concatenate([old.charge, new.charge])
        got, temp, _ = oracle.ewald_delta(box, pos, q, np.ones(2 * ms), ms, ms, same, cross)
        ref = z["sb_ewald"][k]
        # the reference side is a difference of two totals of magnitude 1e5-1e6: 1e-9 of the totals is the resolution
        tol = 1e-9 * max(1.0, float(np.abs(z["ewald_E"]).max()))
        assert abs(got[0] - ref[0]) <= tol and abs(got[1] - ref[1]) <= tol, (name, k, got, ref)
        act = np.abs(temp.reshape(-1, 2)).sum(axis=1) > 0
        assert np.max(np.abs(temp.reshape(-1, 2)[act] - z["sb_temp"][k].reshape(-1, 2)[act])) < 1e-9


@pytest.mark.parametrize("name", ["A", "E", "B", "C"])
def test_insertion_ewald_delta_matches_reference_totals(oracle, name):
    """INSERTION: Fourier delta minus the rigid exclusion constants (Ewald_Energy_Functions.h:547-576) = Ewald_Total with the
    molecule - Ewald_Total without it, both by the reference"""
    box, ff, s, z = load_config(name)
    comp = int(z["comp"]); ms = int(s.molsize[comp]); o = int(s.offsets[comp])
    q = s.charge[o:o + ms]
    got, temp, _ = oracle.ewald_delta(box, z["ins_pos"], q, np.ones(ms), 0, ms, z["sf_ads"], z["sf_fw"])
    same = got[0] - float(z["excl"][0]) - float(z["excl"][1])
    tol = 1e-9 * max(1.0, float(np.abs(z["ewald_E"]).max()))
    assert abs(same - z["ins_ewald"][0]) <= tol and abs(got[1] - z["ins_ewald"][1]) <= tol
    act = np.abs(temp.reshape(-1, 2)).sum(axis=1) > 0
    assert np.max(np.abs(temp.reshape(-1, 2)[act] - z["ins_temp"].reshape(-1, 2)[act])) < 1e-9


def test_rng_stream_seed0(oracle):
    u = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "rng_seed0.npz"))["u"]
    assert (oracle.uniform_stream(0, len(u)) == u).all()


def test_ewald_setup_matches_published_parameters(oracle):
    # config A alpha = 0.225533, kmax (8,11,9), nvec 3933, ReciprocalCutOff 133.4025;
    # config B alpha = 0.265058, (11,11,7), 4140 (Examples/CO2-MFI/output.txt:80)
    box, ff, s, z = load_config("A")
    b = oracle.ewald_setup(box, 14.0, 1e-6)
    assert abs(b.alpha - 0.225533) < 5e-7 and b.kmax == (8, 11, 9) and b.nvec == 3933 and abs(b.recip_cutoff - 133.4025) < 1e-9
    box, ff, s, z = load_config("B")
    b = oracle.ewald_setup(box, 12.0, 1e-6)
    assert abs(b.alpha - 0.265058) < 5e-7 and b.kmax == (11, 11, 7) and b.nvec == 4140


# NIST SPC/E reference calculations in non-cuboid cells, as tabulated in the reference's
# Examples/Reference_NIST_SPCE/readme.md:7-17 (Kelvin): E_disp, E_tail, E_real, number of wave vectors, E_fourier,
# E_self + E_intra, E_total
NIST_SPCE = {
    1: (111992.0, -4109.19, -727219.0, 831, 44677.0, -146595.0, -721254.0),
    2: (43286.0, -2105.61, -476902.0, 1068, 44409.4, -109946.0, -501259.0),
    3: (14403.3, -1027.3, -297129.0, 838, 28897.4, -73298.0, -328153.0),
    4: (25025.1, -163.091, -171462.0, 1028, 22337.2, -36648.0, -160912.0),
}


@pytest.mark.parametrize("b", [1, 2, 3, 4])
def test_nist_spce_known_answers(oracle, b):
    """The reference's own known-answer test for this path: total VDW / real / Fourier / self+intra / tail of 100-400
    SPC/E waters in triclinic boxes, LAMMPS-style Ewald set-up.  The oracle must reproduce the reference's printed
    energies to every printed decimal, and NIST's table to its 6 significant figures."""
    from tests.conftest import load_nist, ENERGY_TO_KELVIN as K
    box, ff, s, ref = load_nist(b)
    v = oracle.total_vdw_real(box, ff, s)                 # HHv, HHr, HGv, HGr, GGv, GGr
    E, sa, sf = oracle.ewald_total(box, s)                # GG (Fourier - self - intra), HH, HG
    tail = oracle.tail_total(ff, pseudo_atom_counts(s, ff.ntypes), box.volume)
    assert v[0] == v[1] == v[2] == v[3] == 0.0 and E[1] == E[2] == 0.0
    # the reference's printed values (5 decimals)
    assert abs(v[4] - ref["vdw_gg"]) < 1e-5 and abs(v[5] - ref["real_gg"]) < 1e-5
    assert abs(E[0] - ref["ewald_gg"]) < 1e-5 and abs(tail - ref["tail"]) < 1e-5
    assert abs(v.sum() + E.sum() + tail - ref["total"]) < 2e-5
    # NIST's table
    disp, etail, real, nk, fourier, self_intra, total = NIST_SPCE[b]
    assert int((np.abs(sa.reshape(-1, 2)).sum(axis=1) > 0).sum()) == nk
    assert abs(v[4] * K - disp) <= 1e-5 * abs(disp) and abs(tail * K - etail) <= 1e-5 * abs(etail)
    assert abs(v[5] * K - real) <= 1e-5 * abs(real)
    assert abs(ref["fourier_gg"] * K - fourier) <= 5e-5 * abs(fourier)
    assert abs((E[0] - ref["fourier_gg"]) * K - self_intra) <= 5e-5 * abs(self_intra)
    assert abs((v.sum() + E.sum() + tail) * K - total) <= 1e-5 * abs(total)


def test_scale_positions_properties(oracle):
    """ScalePositions restatement (mc_box.h:18-64): the identity at scale 1, first atoms scale with the box, molecules stay rigid,
    the framework does not move"""
    from tests.conftest import load_config
    box, ff, s, z = load_config("B")
    comp = 1; ms = 3; o = int(s.offsets[comp]); nm = int(s.natoms[comp]) // ms
    assert np.array_equal(oracle.scale_positions(box, s, 1.0)[:int(s.natoms[0])], s.pos[:int(s.natoms[0])])
    assert np.max(np.abs(oracle.scale_positions(box, s, 1.0) - s.pos)) < 1e-12
    p = oracle.scale_positions(box, s, 1.05)
    assert np.array_equal(p[:int(s.natoms[0])], s.pos[:int(s.natoms[0])])
    old = s.pos[o:o + nm * ms].reshape(nm, ms, 3); new = p[o:o + nm * ms].reshape(nm, ms, 3)
    assert np.max(np.abs(new[:, 0] - 1.05 * old[:, 0])) < 1e-12
    d_old = np.linalg.norm(old[:, 1:] - old[:, :1], axis=2); d_new = np.linalg.norm(new[:, 1:] - new[:, :1], axis=2)
    assert np.max(np.abs(d_old - d_new)) < 1e-12 and d_old.max() < 3.0


def _with(s, comp, scale_coul=None, extra_pos=None, extra_scoul=1.0):
    """copy of System s: new scale_coul per atom and / or one more molecule of `comp` (positions extra_pos, the template's charges and types)"""
    from graspa_b200.types import System
    ms = int(s.molsize[comp]); o = int(s.offsets[comp]); nl = int(s.natoms[comp])
    scc = (s.scale_coul if scale_coul is None else scale_coul).copy()
    if extra_pos is None:
        return System(s.nhost, s.natoms.copy(), s.molsize.copy(), s.pos.copy(), s.charge.copy(), s.type.copy(), s.molid.copy(), s.scale.copy(), scc, alloc=s.alloc.copy())
    natoms = s.natoms.copy(); natoms[comp] += ms
    alloc = np.maximum(s.alloc, natoms)
    chunks = {k: [] for k in ("pos", "charge", "type", "molid", "scale", "scoul")}
    for c in range(s.ncomp):
        a = int(s.offsets[c]); n = int(s.alloc[c])
        cols = dict(pos=s.pos[a:a + n], charge=s.charge[a:a + n], type=s.type[a:a + n], molid=s.molid[a:a + n], scale=s.scale[a:a + n], scoul=scc[a:a + n])
        if c == comp:
            live = {k: v[:nl] for k, v in cols.items()}
            add = dict(pos=np.asarray(extra_pos).reshape(ms, 3), charge=s.charge[o:o + ms], type=s.type[o:o + ms], molid=np.full(ms, nl // ms),
                       scale=np.ones(ms), scoul=np.full(ms, extra_scoul))
            pad = int(alloc[c]) - nl - ms
            cols = {k: np.concatenate([live[k], add[k], np.zeros((pad,) + live[k].shape[1:], dtype=live[k].dtype)]) for k in cols}
        for k in chunks:
            chunks[k].append(cols[k])
    cat = {k: np.concatenate(v) for k, v in chunks.items()}
    return System(s.nhost, natoms, s.molsize.copy(), cat["pos"], cat["charge"], cat["type"], cat["molid"], cat["scale"], cat["scoul"], alloc=alloc)


def test_cbcf_two_step_fourier_deltas_on_the_temp_vector_match_reference_totals(oracle):
    """The two-step CB/CFC moves keep their intermediate Fourier state in tempEik (UseTempVector, Ewald_Energy_Functions.h:362-377, :510-517,
    :713-716): the second step's delta is taken against the FIRST step's new structure factors, not against the stored ones.  Restated with
    orc_ewald_delta and pinned to the REFERENCE's Ewald_Total of the three states (config B, CO2 in MFI):
      insertion: molecule m at lambda_c = 0.4^5 -> 1, then a new molecule at lambda_c = 0.7^5;
      deletion : that new molecule removed, then molecule k from 1 -> 0.3^5 on top of it."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built")
    box, ff, s, z = load_config("B")
    comp = int(z["comp"]); ms = int(s.molsize[comp]); o = int(s.offsets[comp]); q = s.charge[o:o + ms]
    excl = float(z["excl"][0]) + float(z["excl"][1])
    m, k = 2, 4
    pm = s.pos[o + m * ms:o + (m + 1) * ms]; pk = s.pos[o + k * ms:o + (k + 1) * ms]
    sc0, sc2, sc3 = 0.4 ** 5, 0.7 ** 5, 0.3 ** 5
    scc = s.scale_coul.copy(); scc[o + m * ms:o + (m + 1) * ms] = sc0
    s1 = _with(s, comp, scale_coul=scc)                                    # before: molecule m is the fractional one
    E1, sa1, sf1 = oracle.ref_ewald_total(box, s1)
    # the reference side is a difference of totals of magnitude 1.5e7 (GG includes HH, ewald_preparation.h:174): 1e-12 of that is a few times
    # its double-precision resolution and 200 times smaller than what using the wrong vector would change (checked below)
    tol = 1e-12 * max(1.0, float(np.abs(E1).max()))
    same = lambda E: E[0]                                                  # guest-guest Fourier total incl. exclusions (ewald_E[0])
    cross = lambda E: E[2] if len(E) > 2 else 0.0
    # ---- insertion, step 1 (lambda change from the stored vectors) and step 2 (growth, continuing from the temp vector of step 1)
    d1, t1, _ = oracle.ewald_delta(box, np.concatenate([pm, pm]), np.concatenate([q, q]), np.concatenate([np.full(ms, sc0), np.ones(ms)]), ms, ms, sa1, sf1)
    d2, t2, _ = oracle.ewald_delta(box, z["ins_pos"], q, np.full(ms, sc2), 0, ms, t1, sf1)
    s2 = _with(_with(s, comp), comp, extra_pos=z["ins_pos"], extra_scoul=sc2)       # after: m whole again, the new molecule fractional
    E2, sa2, sf2 = oracle.ref_ewald_total(box, s2)
    got_same = (d1[0] - excl * (1.0 - sc0 ** 2)) + (d2[0] - excl * sc2 ** 2)
    assert abs(got_same - (same(E2) - same(E1))) <= tol, (got_same, same(E2) - same(E1))
    assert abs((d1[1] + d2[1]) - (cross(E2) - cross(E1))) <= tol
    act = np.abs(t2.reshape(-1, 2)).sum(axis=1) > 0
    assert np.max(np.abs(t2.reshape(-1, 2)[act] - sa2.reshape(-1, 2)[act])) < 1e-9          # what the acceptance swaps in
    # the second step taken against the STORED vectors instead (what Insertion_Body's INSERTION call does in the reference snapshot) differs
    d2_wrong, _, _ = oracle.ewald_delta(box, z["ins_pos"], q, np.full(ms, sc2), 0, ms, sa1, sf1)
    print("second step against the stored vectors instead of the temp vector: off by", d2_wrong[0] - d2[0])
    assert abs(d2_wrong[0] - d2[0]) > 100 * tol
    # ---- deletion of the fractional molecule (step 1 into temp), then molecule k from 1 to sc3 on top of it (UseTempVector)
    d3, t3, _ = oracle.ewald_delta(box, z["ins_pos"], q, np.full(ms, sc2), ms, 0, sa2, sf2)
    d4, t4, _ = oracle.ewald_delta(box, np.concatenate([pk, pk]), np.concatenate([q, q]), np.concatenate([np.ones(ms), np.full(ms, sc3)]), ms, ms, t3, sf2)
    scc3 = s.scale_coul.copy(); scc3[o + k * ms:o + (k + 1) * ms] = sc3
    s3 = _with(s, comp, scale_coul=scc3)
    E3, sa3, _ = oracle.ref_ewald_total(box, s3)
    got_same = (d3[0] - excl * (0.0 - sc2 ** 2)) + (d4[0] - excl * (sc3 ** 2 - 1.0))
    assert abs(got_same - (same(E3) - same(E2))) <= tol, (got_same, same(E3) - same(E2))
    assert abs((d3[1] + d4[1]) - (cross(E3) - cross(E2))) <= tol
    act = np.abs(t4.reshape(-1, 2)).sum(axis=1) > 0
    assert np.max(np.abs(t4.reshape(-1, 2)[act] - sa3.reshape(-1, 2)[act])) < 1e-9
