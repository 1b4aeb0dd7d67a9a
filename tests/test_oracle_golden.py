"""The plain-C oracle against the golden vectors generated from the reference's own routines
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from graspa_b200.types import TrialAtoms, species_counts, pseudo_atom_counts
from tests.conftest import load_config, rel_err

CONFIGS = ["A", "E", "B", "D"]


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


@pytest.mark.parametrize("name", ["A", "E", "B"])
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
